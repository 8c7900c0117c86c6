"""The experiment of the reference's README (``legacy/examples/llama2_4D_finetune/exp.py`` + ``figures/``): run the same finetune
single-device and under several DP x TP layouts, overlay the loss curves.

    python examples/llama_4D_finetune/exp.py --layouts 1x1 2x2 4x1 1x4 --max_iters 30 [--out_dir /tmp/llama_exp]

Writes one log per layout, ``curves.csv`` (iteration, one loss column per layout) and — when matplotlib is importable —
``curves.png``; prints the largest deviation of every layout from the single-device curve and exits non-zero if it exceeds
``--tol``.  Runs on CPU (gloo) when no GPU is present.
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def parse_log(path):
    """``{"train": {iter: loss}, "eval": {iter: (train loss, val loss)}}`` from a llama_train.py log."""
    tr, ev = {}, {}
    with open(path) as f:
        for line in f:
            m = re.match(r"iter (\d+): loss ([0-9.]+)", line)
            if m:
                tr[int(m.group(1))] = float(m.group(2))
            m = re.match(r"eval (\d+): train loss ([0-9.]+), val loss ([0-9.]+)", line)
            if m:
                ev[int(m.group(1))] = (float(m.group(2)), float(m.group(3)))
    return {"train": tr, "eval": ev}


def run_layout(dp, tp, a, port):
    log = os.path.join(a.out_dir, f"dp{dp}_tp{tp}.log")
    if os.path.exists(log):
        os.remove(log)
    common = ["--dp", str(dp), "--tp", str(tp), "--max_iters", str(a.max_iters), "--config", a.config, "--bsz", str(a.bsz), "--seqlen", str(a.seqlen),
              "--eval_interval", str(a.eval_interval), "--log_file", log] + a.extra
    script = os.path.join(HERE, "llama_train.py")
    n = dp * tp
    if n == 1:
        cmd = [sys.executable, script] + common
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(port), script] + common
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", "1")
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=a.timeout)
    if r.returncode != 0:
        raise RuntimeError(f"dp{dp} x tp{tp} failed:\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}")
    return parse_log(log)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", nargs="+", default=["1x1", "2x2"], help="DPxTP, the first one is the baseline")
    ap.add_argument("--max_iters", type=int, default=20)
    ap.add_argument("--config", default="tiny")
    ap.add_argument("--bsz", type=int, default=8)
    ap.add_argument("--seqlen", type=int, default=32)
    ap.add_argument("--eval_interval", type=int, default=10)
    ap.add_argument("--out_dir", default=os.path.join(HERE, "exp_out"))
    ap.add_argument("--tol", type=float, default=5e-3)
    ap.add_argument("--timeout", type=int, default=1800)
    ap.add_argument("--port", type=int, default=29840)
    ap.add_argument("extra", nargs=argparse.REMAINDER, help="passed through to llama_train.py after '--'")
    a = ap.parse_args()
    a.extra = [x for x in a.extra if x != "--"]
    os.makedirs(a.out_dir, exist_ok=True)
    curves = {}
    for k, lay in enumerate(a.layouts):
        dp, tp = (int(x) for x in lay.lower().split("x"))
        curves[lay] = run_layout(dp, tp, a, a.port + k)
        tr = curves[lay]["train"]
        print(f"[{lay}] {len(tr)} iterations, loss {tr[min(tr)]:.4f} -> {tr[max(tr)]:.4f}", flush=True)
    base = a.layouts[0]
    its = sorted(curves[base]["train"])
    with open(os.path.join(a.out_dir, "curves.csv"), "w") as f:
        f.write("iter," + ",".join(a.layouts) + "\n")
        for it in its:
            f.write(f"{it}," + ",".join(f"{curves[l]['train'].get(it, float('nan')):.6f}" for l in a.layouts) + "\n")
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt

        for lay in a.layouts:
            plt.plot(its, [curves[lay]["train"][i] for i in its], label=f"dp x tp = {lay}", linewidth=1.2)
        plt.xlabel("iteration"), plt.ylabel("training loss"), plt.legend()
        plt.savefig(os.path.join(a.out_dir, "curves.png"), dpi=120)
    except ImportError:
        pass
    worst = 0.0
    for lay in a.layouts[1:]:
        d = max(abs(curves[lay]["train"][i] - curves[base]["train"][i]) for i in its)
        dv = max((abs(curves[lay]["eval"][i][1] - curves[base]["eval"][i][1]) for i in curves[base]["eval"]), default=0.0)
        print(f"max |{lay} - {base}| over {len(its)} iterations: train {d:.2e}, val {dv:.2e}")
        worst = max(worst, d, dv)
    if worst > a.tol:
        print(f"FAILED: curves differ by {worst:.2e} > {a.tol}")
        sys.exit(1)
    print("loss curves agree")


if __name__ == "__main__":
    main()
