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


def test_llama_4d_finetune_hf_model_curves_and_resume(tmp_path):
    """An UNMODIFIED HuggingFace LlamaForCausalLM under DP x TP + SP (DModule plan, DDP, DistributedOptimizer) on the token-bin data
    pipeline: ``exp.py`` overlays dp2 x tp2 and dp1 x tp4 (TP only) on the single-device curve; a checkpoint written at iteration 4
    resumes onto the same curve (legacy ``examples/llama2_4D_finetune`` README experiment)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    out = str(tmp_path / "exp")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples/llama_4D_finetune/exp.py"), "--layouts", "1x1", "2x2", "--max_iters", "8", "--eval_interval", "4",
                        "--out_dir", out, "--port", "29710"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "loss curves agree" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    sys.path.insert(0, os.path.join(ROOT, "examples", "llama_4D_finetune"))
    from exp import parse_log

    full = parse_log(os.path.join(out, "dp2_tp2.log"))["train"]
    assert full[7] < full[0]
    ck = str(tmp_path / "ck")
    _torchrun(4, "examples/llama_4D_finetune/llama_train.py", "--dp", "2", "--tp", "2", "--max_iters", "4", "--save_interval", "4", "--ckpt_dir", ck, "--async_checkpoint", port=29712)
    log = str(tmp_path / "resumed.log")
    _torchrun(4, "examples/llama_4D_finetune/llama_train.py", "--dp", "2", "--tp", "2", "--max_iters", "8", "--ckpt_dir", ck, "--resume", "--log_file", log, port=29713)
    resumed = parse_log(log)["train"]
    assert sorted(resumed) == [4, 5, 6, 7], resumed
    for k, v in resumed.items():
        assert abs(v - full[k]) < 2e-3, (k, v, full[k])


def test_mixtral_ep_training_hf_model_matches_single_device():
    """Unmodified HF MixtralForCausalLM, experts hijacked onto EP=4 (one expert per rank per layer, so ranks regularly receive no
    token for a layer), MoEOptimizer: the loss curve equals the unparallelised twin on the global batch."""
    out = _torchrun(4, "examples/mixtral_EP_training/mixtral_train.py", "--max_iters", "8", "--compare-single", port=29715)
    assert "loss curves agree" in out, out[-2000:]
