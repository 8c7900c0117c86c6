"""End-to-end example scripts as tests (legacy ``examples/*/README`` experiments: the parallel run's loss curve lies on the
single-device one)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, *args, port=29700, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script), *args]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


def test_mixtral_4d_training_matches_single_device_and_resumes(tmp_path):
    """Mixtral (sparse MoE) under DP x TP(+SP) + DistributedOptimizer: loss curve equal to the unparallelised twin, loss goes down on
    the structured corpus; a checkpoint written at iteration 4 resumes onto the same curve."""
    out = _torchrun(4, "examples/mixtral_4D_training/mixtral_train.py", "--dp", "2", "--tp", "2", "--max_iters", "8", "--compare-single", port=29701)
    assert "loss curves agree" in out, out[-2000:]
    full = {l.split(":")[0]: float(l.split("loss")[1].split()[0]) for l in out.splitlines() if l.startswith("iter ")}
    ck = str(tmp_path / "ck")
    _torchrun(4, "examples/mixtral_4D_training/mixtral_train.py", "--dp", "2", "--tp", "2", "--max_iters", "4", "--save_interval", "4", "--ckpt_dir", ck, port=29702)
    out = _torchrun(4, "examples/mixtral_4D_training/mixtral_train.py", "--dp", "2", "--tp", "2", "--max_iters", "8", "--ckpt_dir", ck, "--resume", port=29703)
    resumed = {l.split(":")[0]: float(l.split("loss")[1].split()[0]) for l in out.splitlines() if l.startswith("iter ")}
    assert sorted(resumed) == ["iter 4", "iter 5", "iter 6", "iter 7"], resumed
    for k, v in resumed.items():
        assert abs(v - full[k]) < 2e-3, (k, v, full[k])
