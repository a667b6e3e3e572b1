"""GPU: a training step replayed as one CUDA graph (core/graph_step.py) is the step the eager trainer runs -- same
random numbers (bottleneck eps, fused dropout, the samples of marginal_kl), same RAdam schedule, same losses."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=7295, dropout=True):
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    from sparse_vae_b200.data_parallel import GradientAllReducer
    dev = torch.device('cuda')
    torch.manual_seed(seed)
    hp = sv.TransformerVAEHparams(d_model=512, num_layers=4, num_heads=8, latent_depth=32)
    model = sv.TransformerVAE(to_attrdict(hp)).to(dev)
    model.initialize_weights()
    model.train()
    model.validate_posterior = False
    if not dropout:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    (opt,), (cfg,) = model.configure_optimizers(tokens_per_batch=4 * 512, accumulate_grad_batches=1)
    return model, opt, cfg['scheduler'], GradientAllReducer(model)


def _batches(n):
    from sparse_vae_b200.synthetic import synthetic_tokens, to_device
    dev = torch.device('cuda')
    return [to_device(synthetic_tokens(4, 512, seed=100 + i, lengths=[512, 480, 512, 301]), dev, non_blocking=False) for i in range(n)]


def test_graphed_step_matches_eager_step():
    from sparse_vae_b200.core.graph_step import GraphedTrainStep
    batches = _batches(7)
    losses = {}
    params = {}
    for mode in ('eager', 'graph'):
        # dropout off: ATen's random ops (encoder dropout) would otherwise sit BETWEEN the library's draws in program order,
        # and a replay gives ATen the offsets after the library's -- valid, but not the eager step's numbers
        model, opt, sched, reducer = _setup(dropout=False)
        torch.manual_seed(11)                       # the step's random streams start from the same generator state
        step = GraphedTrainStep(model, opt, sched, reducer, torch.bfloat16, warmup=2)
        fn = step.eager if mode == 'eager' else step
        losses[mode] = [float(fn(b)) for b in batches]
        params[mode] = torch.cat([p.detach().flatten() for p in model.parameters()])
        if mode == 'graph':
            assert step.graph is not None and step.calls == len(batches)
            assert opt.param_groups[0]['step'] == len(batches) + 1
            assert model.global_step == len(batches)
    for a, b in zip(losses['eager'], losses['graph']):
        assert abs(a - b) <= 1e-5 * abs(a), (losses['eager'], losses['graph'])
    rel = (params['eager'] - params['graph']).abs().max().item() / params['eager'].abs().max().item()
    assert rel <= 1e-4, rel


def test_graph_replays_draw_fresh_random_numbers():
    """Two replays on the SAME batch must differ (new eps, new dropout masks) and leave the generator where eager would."""
    from sparse_vae_b200.core.graph_step import GraphedTrainStep
    model, opt, sched, reducer = _setup()
    for g in opt.param_groups:
        g['lr'] = 0.0                               # frozen weights: differences can only come from the random draws
    sched = None
    step = GraphedTrainStep(model, opt, sched, reducer, torch.bfloat16, warmup=2)
    b = _batches(1)[0]
    vals = [float(step(b)) for _ in range(6)]
    assert len(set(vals[2:])) == len(vals[2:]), vals
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    before = gen.get_offset()
    step(b)
    assert gen.get_offset() - before >= step.philox.delta > 0
