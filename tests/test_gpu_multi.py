"""Several GPUs (NCCL over NVLink): results must not depend on the number of GPUs or on the grid.
Needs >= 2 GPUs on the box (gpurun --gpus N); skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _launch(world, args, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")] + [str(a) for a in args]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0 and "MULTI-GPU PARITY OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("world,grid", [(2, "-"), (2, "2x1"), (4, "-"), (4, "1x4"), (8, "-")])
def test_multi_gpu_parity_fixture(world, grid):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, ["reads_fa", 17, 2, 8, grid], 29611 + world)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity_synthetic_hifi(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, ["synth:300000,500,12000,0.01", 31, 2, 4, "-", 16], 29631 + world)


@pytest.mark.parametrize("world,grid", [(2, "-"), (4, "2x2")])
def test_multi_gpu_hash_path_forced_k31(world, grid):
    """k = 31 counts through super-k-mers on several GPUs too (every GPU parses all reads, counts its own buckets);
    the two-level hash partition with its all-to-all must still give the same bits."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, ["synth:300000,500,12000,0.01", 31, 2, 4, grid, 16], 29651 + world, env={"ELBA_FE_COUNT_PATH": "hash"})


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_superkmer_other_geometries(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, ["synth:200000,300,9000,0.02", 25, 2, 6, "-"], 29671 + world)
    _launch(world, ["reads_fa", 21, 2, 8, "-"], 29681 + world)


@pytest.mark.parametrize("world,grid", [(2, "-"), (4, "2x2")])
def test_multi_gpu_nonzero_first_read_id(world, grid):
    """The global read ids need not start at 0 (elba_fe_upload_reads takes any offset for rank 0)."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, ["synth:200000,300,9000,0.02", 31, 2, 4, grid, 0, 1000], 29691 + world)
    _launch(world, ["reads_fa", 17, 2, 8, grid, 0, 77], 29695 + world)


@pytest.mark.parametrize("world", [2])
def test_multi_gpu_without_peer_memory(world):
    """ELBA_FE_P2P=0: the super-k-mer path needs peer access between the GPUs; without it every k falls back to the hash path
    with its NCCL all-to-all.  Same bits."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, ["synth:300000,500,12000,0.01", 31, 2, 4, "-", 16], 29699 + world, env={"ELBA_FE_P2P": "0"})


@pytest.mark.parametrize("world,grid", [(2, "-"), (2, "2x1"), (4, "2x2"), (8, "-")])
def test_multi_gpu_xdrop_alignment(world, grid):
    """elba_fe_align on several GPUs: every rank aligns the nonzeros of its block of B (reads of the other ranks all-gathered
    over NVLink); square grids select pairs by the reference's block-local rule (src/PairwiseAlignment.cpp:52), the others by
    the global upper triangle.  Every pair exactly once, 13 fields bit-exact against the oracle, nonzero first read id."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, ["synth:120000,260,7000,0.05", 21, 2, 6, grid, 0, 0, "align"], 29711 + world)
    _launch(world, ["reads_fa", 17, 2, 8, grid, 0, 41, "align"], 29721 + world)
