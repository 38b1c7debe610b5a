"""pipeline.HostPipeline: batches in pinned host memory, copies overlapped with the compute of the neighbouring batches."""
import pytest
import torch

import common as C


@pytest.mark.gpu
def test_host_pipeline_returns_every_batch_in_order():
    from idccrn_b200.pipeline import HostPipeline
    enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 3, "cuda")
    x, eps = C.vae_inputs(3, 4000, 1, 1, 3, "cuda")

    def enhance(xd):
        return C.run_vae(enc, dec, xd, eps, "skip_prepare")["recon_sig"]
    xs = [(x * (0.5 + 0.25 * i)).cpu().pin_memory() for i in range(5)]
    want = [enhance(t.cuda()).cpu() for t in xs]
    hp = HostPipeline(enhance, "cuda", depth=2)
    outs = [torch.empty_like(want[0]).pin_memory() for _ in xs]
    for t, o in zip(xs, outs):
        hp.submit(t, o)
    hp.join()
    torch.cuda.synchronize()
    for w, o in zip(want, outs):
        assert torch.equal(w, o)
