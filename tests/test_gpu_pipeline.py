"""HostPipeline (vqvae_b200/pipeline.py): the streaming host-buffer front end returns, batch for
batch, exactly what a direct ``model(x.cuda())`` returns (same kernels, same order)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(prec):
    import vqvae_b200
    from oracle import weights
    sd = weights.make_state_dict(128, 32, 2, 512, 64, seed=5)
    m = vqvae_b200.VQVAE(128, 32, 2, 512, 64, 0.25)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    vqvae_b200.set_precision(prec)
    return m.cuda().eval()


@pytest.mark.parametrize("use_graph", [True, False])
@pytest.mark.parametrize("prec", ["tf32", "fp32"])
def test_pipeline_matches_direct_forward(use_graph, prec):
    import vqvae_b200
    try:
        m = _model(prec)
        rng = np.random.default_rng(11)
        batches = [torch.from_numpy(rng.standard_normal((16, 3, 32, 32)).astype(np.float32)).pin_memory()
                   for _ in range(7)]
        want = []
        with torch.no_grad():
            for x in batches:
                loss, x_hat, perp = m(x.cuda())
                want.append((float(loss), x_hat.cpu().clone(), float(perp)))
        pipe = vqvae_b200.HostPipeline(m, (16, 3, 32, 32), depth=3, use_graph=use_graph)
        got = []
        n = pipe.run(batches, lambda r: got.append((r.index, float(r.loss), r.x_hat.clone(), float(r.perplexity))))
        assert n == 7 and [g[0] for g in got] == list(range(7))
        for (l0, xh0, p0), (_, l1, xh1, p1) in zip(want, got):
            assert l0 == l1 and p0 == p1
            assert torch.equal(xh0, xh1)
    finally:
        vqvae_b200.set_precision("fp32")


def test_pipeline_rejects_wrong_input():
    import vqvae_b200
    m = _model("fp32")
    pipe = vqvae_b200.HostPipeline(m, (4, 3, 32, 32), depth=2, use_graph=False)
    with pytest.raises(ValueError):
        pipe.push(torch.zeros(4, 3, 16, 16))
    with pytest.raises(ValueError):
        pipe.push(torch.zeros(4, 3, 32, 32, dtype=torch.float64))
    assert pipe.drain() == []
