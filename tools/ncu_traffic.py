"""profiles/traffic.json from an `ncu --set full` raw CSV export of ONE pass of a workload: DRAM bytes (read + write) per
launch of every kernel of ours, and the sum over the counting chain.  usage: ncu_traffic.py raw.csv workload source-note"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
workload, note = sys.argv[2], sys.argv[3]
h, units = rows[0], rows[1]
idx = {n: i for i, n in enumerate(h)}
def gb(r, n):
    v = float(r[idx[n]].replace(',', '')); u = units[idx[n]]
    if v != v:          # ncu prints nan for a launch too short to sample
        v = 0.0
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
kern = {}
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('elba::', '').split('<')[0]
    k = kern.setdefault(name, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "ms": 0.0})
    k["launches"] += 1
    k["dram_read"] += gb(r, 'dram__bytes_read.sum'); k["dram_write"] += gb(r, 'dram__bytes_write.sum')
    k["ms"] += float(r[idx['gpu__time_duration.sum']].replace(',', '')) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[idx['gpu__time_duration.sum']]]
# the counting chain = what the reference does in get_kmer_count_map_keys / values (the radix sort of the reliable k-mers is CUB's and
# shares its kernel names with the sorts of build_A: it is left out of this sum and listed by name)
chain = ("k_skm_scatter", "k_skm_forward", "k_skm_plan", "k_skm_count4", "k_skm4_count_global", "k_skm4_collect_global", "k_skm4_emit_global", "k_table_clear",
         "k_iota_u32", "k_rank_finish", "k_rank_global", "k_seed_keys", "k_seed_route", "k_route_unpack",
         "k_unmix", "k_lookup_build", "k_resolve", "k_scatter1", "k_scatter2", "k_count_buckets", "k_probe_filter", "k_count_array", "k_table_collect")
spgemm = ("k_sp2_expand", "k_sp2_warp", "k_sp2_block", "k_sp2_global")
path = os.path.join(ROOT, "profiles", "traffic.json")
try:
    out = json.load(open(path))
except Exception:
    out = {}
out[workload] = {"source": note, "chain_dram_bytes": sum(v["dram_read"] + v["dram_write"] for n, v in kern.items() if n in chain),
                 "spgemm_dram_bytes": sum(v["dram_read"] + v["dram_write"] for n, v in kern.items() if n in spgemm),
                 "kernels": {n: {a: (round(b, 3) if a == "ms" else int(b)) for a, b in v.items()} for n, v in kern.items()}}
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out[workload], indent=1))
