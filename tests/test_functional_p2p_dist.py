"""Functional send / recv (native piece N-C4: ``legacy/patches/patched_pytorch_v2.2.1_rc3.patch:1595-1738`` adds them to torch's
functional collectives): custom ops with autograd across ranks and fake kernels for tracing."""
import torch

from common import run_distributed


def _fp2p(rank, world):
    from vescale_b200.comm import functional_p2p as fp

    w = torch.full((3, 4), float(rank + 1), requires_grad=True)
    x = torch.arange(12.0).view(3, 4)
    if rank == 0:
        tok = fp.send(x * w, 1)  # stage 0: y = x * w0 goes to rank 1
        tok.sum().backward()  # its backward receives dL/dy from rank 1
        torch.testing.assert_close(w.grad, x * 2.0)  # dL/dy = w1 = 2
        fp.isend(torch.ones(2), 1).wait()
    else:
        y = fp.recv((3, 4), torch.float32, 0, requires_grad=True)
        (y * w).sum().backward()  # sends dL/dy = w1 back to rank 0
        torch.testing.assert_close(w.grad, x * 1.0)  # y = x * w0, w0 = 1
        assert torch.equal(fp.irecv((2,), torch.float32, 0).wait(), torch.ones(2))
    from torch._subclasses.fake_tensor import FakeTensorMode

    with FakeTensorMode():  # shape propagation without communication (what a tracer sees)
        t = torch.ops.vescale_b200.p2p_recv(torch.empty(0), [5, 6], 0, 0)
        assert t.shape == (5, 6)
        assert torch.ops.vescale_b200.p2p_send(torch.empty(3, 2), 1, 0).numel() == 0


def test_functional_send_recv():
    run_distributed(_fp2p, 2)
