/* TEST INFRASTRUCTURE: the C-ABI entry points the shim's translation unit references but the transitive-reduction harness never calls */
#define STUB(name) int name() { return -1; }
STUB(elba_fe_build_A) STUB(elba_fe_comm_get_id) STUB(elba_fe_comm_init) STUB(elba_fe_count) STUB(elba_fe_get_A) STUB(elba_fe_get_AT) STUB(elba_fe_get_B_triples)
STUB(elba_fe_get_kmers) STUB(elba_fe_get_reads) STUB(elba_fe_ingest_fasta) STUB(elba_fe_reads_size) STUB(elba_fe_sizes) STUB(elba_fe_spgemm) STUB(elba_fe_upload_reads)
