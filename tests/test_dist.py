"""Multi-rank path: world_size-2/3 gloo runs on CPU for the host-side logic (partition plan, halo/send lists,
exchange protocol), and the real thing on GPUs (`-m gpu`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def launch(mode, world, port, timeout=600, extra_env=None):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.parametrize("world", [2, 3])
def test_partition_plan_and_exchange_on_cpu(world):
    out = launch("cpu", world, 29600 + world)
    assert out.returncode == 0 and "DIST_OK cpu" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_fem_plan_from_mesh_on_cpu(world):
    out = launch("cpu-fem", world, 29650 + world)
    assert out.returncode == 0 and "DIST_OK cpu-fem" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_distributed_solve_on_gpus(nbgpu_lib, world):
    """One rank per GPU when the box has several; on a single-GPU box both ranks share GPU 0 (CUDA IPC works
    within one device; the kernels of the two processes are time-sliced, so this only checks correctness)."""
    out = launch("gpu", world, 29700 + world, timeout=900, extra_env={"NBGPU_DIST_TIMEOUT_MS": "60000"})
    assert out.returncode == 0 and "DIST_OK gpu" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_distributed_fem_assembles_rank_local_rows_on_the_device(nbgpu_lib, world):
    """nbgpu_dist_fem_*: per-rank sub-mesh assembly straight into the rank-local block (no host round trip of K),
    bit-identical rows on all four reference fixtures (slabs and unstructured node ranges), distributed solve."""
    out = launch("gpu-fem", world, 29800 + world, timeout=900, extra_env={"NBGPU_DIST_TIMEOUT_MS": "60000"})
    assert out.returncode == 0 and "DIST_OK gpu-fem" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
def test_every_visible_gpu_at_bench_size_with_parity(nbgpu_lib):
    """All GPUs of the box, ~1 M dof per rank (the bench workload): `bench.py --gpus N` must come back with a green
    parity block -- rank-local rows bit-identical to the single-GPU matrix and to the CPU reference's, 50 iterations
    within 1e-12 of both, full-solve iteration count within +-2 % of the single-GPU count."""
    import json
    n = nbgpu_lib.nbgpu_device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs in the box")
    env = dict(os.environ, NBGPU_DIST_TIMEOUT_MS="60000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29911", os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "1",
           "--warmup", "3", "--no-target"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == n and line["config"]["dof_per_gpu"] >= 1_000_000
    assert line["parity"]["ok"], line["parity"]
