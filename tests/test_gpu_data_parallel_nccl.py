"""GPU, world_size 2, NCCL: two ranks with half the batch each must reproduce the loss and
gradients of one process with the whole batch -- both loss terms are batch means (core/continuous_autoencoder.py:47,
core/language_model.py:161-170), so the averaged per-rank gradient IS the global-batch gradient.  Needs two GPUs on
the box (skipped otherwise; run it with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

B_GLOBAL, L = 4, 512


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _model(dev):
    import sparse_vae_b200 as sv
    from sparse_vae_b200.core.lightning_shim import to_attrdict
    torch.manual_seed(7295)
    hp = sv.TransformerVAEHparams(d_model=512, num_layers=4, num_heads=8, latent_depth=32)
    model = sv.TransformerVAE(to_attrdict(hp)).to(dev)
    model.initialize_weights()
    model.train()
    model.validate_posterior = False
    for m in model.modules():                      # dropout draws differ between the two decompositions of the batch
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, 'dropout_p'):
            m.dropout_p = 0.0
    # the reparameterisation noise is drawn per flat element index of the rank's own [rows, latent] tensor, so two
    # ranks with half the rows each see different eps than one process with all rows: take it out of the comparison
    # with sigma = exp(-15) (the KL term and its gradient stay, they do not depend on eps)
    with torch.no_grad():
        model.q_of_z_given_x.linear.bias[hp.latent_depth:] = -30.0
    return model


def _batch(dev, lo, hi):
    from sparse_vae_b200.synthetic import synthetic_tokens, to_device
    # full-length sequences: the NLL is a mean over a rank's own valid tokens, so the mean of the ranks' losses is the
    # global-batch loss only when every rank holds the same number of them
    host = synthetic_tokens(B_GLOBAL, L, seed=11)
    host = {k: v[lo:hi] for k, v in host.items()}
    return to_device(host, dev, non_blocking=False)


def _step(model, batch, reducer=None):
    # fp32 end to end (the attention runs its exact fp32 kernels): the comparison is then limited by summation order only,
    # not by bf16 GEMMs that pick different tilings for 1024 and 2048 rows
    model.zero_grad(set_to_none=True)
    with torch.autocast('cuda', enabled=False):
        out = model.training_step(batch, 0)
    out['loss'].backward()
    if reducer is not None:
        reducer.finish()
    return out['loss'].detach().float()


def _graph_worker(rank, world, port, out):
    """Two ranks, the split-graph step (graph A, all-reduce behind in-graph events or after the graph, graph B) against
    the eager step on the same data."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sparse_vae_b200.core.graph_step import GraphedTrainStep
    from sparse_vae_b200.data_parallel import GradientAllReducer, init_distributed
    r, local, w = init_distributed('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    per = B_GLOBAL // world
    losses = {}
    for mode in ('eager', 'graph', 'graph_no_overlap'):
        os.environ['SVAE_DP_OVERLAP'] = '0' if mode == 'graph_no_overlap' else '1'
        model = _model(dev)
        (opt,), (cfg,) = model.configure_optimizers(tokens_per_batch=B_GLOBAL * L, accumulate_grad_batches=1)
        reducer = GradientAllReducer(model, bucket_mb=4.0)
        step = GraphedTrainStep(model, opt, cfg['scheduler'], reducer, torch.bfloat16, warmup=2)
        fn = step.eager if mode == 'eager' else step
        torch.manual_seed(99)
        losses[mode] = [float(fn(_batch(dev, rank * per, (rank + 1) * per))) for _ in range(5)]
        if mode != 'eager':
            assert step.graph_b is not None and step.overlap == (mode == 'graph')
            # the overlapped mode packs every bucket inside graph A, each followed by an external event
            assert (len(reducer._capture_order) == len(reducer.buckets)) == step.overlap
        reducer.remove()
    out[rank] = losses
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_split_graph_step_matches_eager():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_graph_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    for rank in range(world):
        for mode in ('graph', 'graph_no_overlap'):
            for a, b in zip(res[rank]['eager'], res[rank][mode]):
                assert abs(a - b) <= 1e-5 * abs(a), (mode, res)


def _worker(rank, world, port, out):
    """Every rank first computes the whole-batch gradients by itself (no reducer), then the two ranks run the
    data-parallel step on half the batch each; the comparison is per parameter, inside the process that holds both."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sparse_vae_b200.data_parallel import GradientAllReducer, init_distributed
    r, local, w = init_distributed('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    model = _model(dev)
    names = {id(p): n for n, p in model.named_parameters()}
    full_loss = _step(model, _batch(dev, 0, B_GLOBAL)).item()
    full = {names[id(p)]: p.grad.detach().clone() for p in model.parameters() if p.grad is not None}
    reducer = GradientAllReducer(model, bucket_mb=4.0)
    per = B_GLOBAL // world
    res = {}
    for it in range(2):                            # step 0 builds the buckets, step 1 runs the overlapped path
        loss = _step(model, _batch(dev, rank * per, (rank + 1) * per), reducer)
        lsum = loss.clone()
        dist.all_reduce(lsum)
        red = {names[id(p)]: p.grad.detach() for p in model.parameters() if p.grad is not None}
        flat = torch.cat([g.flatten().float() for g in red.values()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        worst = max((red[n] - g).norm().item() / (g.norm().item() + 1e-30) for n, g in full.items())
        fnorm = torch.cat([g.flatten() for g in full.values()]).norm().item()
        res[it] = dict(loss=(lsum / world).item(), full_loss=full_loss, gnorm=flat.norm().item(), full_gnorm=fnorm,
                       same=all(torch.equal(gathered[0], t) for t in gathered), same_params=set(red) == set(full),
                       worst_param_rel=worst, buckets=len(reducer.buckets))
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_nccl_step_matches_single_process():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        multi = dict(out)
    for rank in range(world):
        for it in range(2):
            m = multi[rank][it]
            assert m['same'], 'ranks disagree bitwise after the all-reduce'
            assert m['same_params'] and m['buckets'] > 1
            assert abs(m['loss'] - m['full_loss']) <= 1e-5 * abs(m['full_loss']), m
            assert abs(m['gnorm'] - m['full_gnorm']) <= 1e-5 * m['full_gnorm'], m
            assert m['worst_param_rel'] <= 1e-4, m      # every parameter's gradient, relative to its own norm
