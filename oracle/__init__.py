"""CPU oracle for the sparse-vae hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + torch-CPU fp32/fp64) of the reference's
algorithm for the path named in BASELINE.json `north_star`:

  * block-sparse layout           -> oracle/layout.py      (reference core/sparse_attention.py:38-59)
  * sdd -> softmax -> dsd         -> oracle/attention.py   (reference core/sparse_attention.py:75-92,
                                                            triton==1.1.0 blocksparse ops, absent from /root/reference)
  * latent bottleneck + Philox    -> oracle/bottleneck.py  (reference core/conditional_gaussian.py:18-30,
                                                            core/continuous_autoencoder.py:42-52, torch Normal.rsample)
  * whole TransformerVAE step     -> oracle/model.py       (reference transformer_vae.py:42-93 and callees)
  * token-by-token decoding       -> oracle/decoding.py    (reference core/generation.py:40-72, core/attention.py:107-142)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
leg may import it, and only as the checker.  Nothing under `sparse_vae_b200/` imports it; the
product path fails loudly when the CUDA extension is missing.

Pinning: the reference ships NO tests, fixtures or golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, run in the build container
behind import stubs (oracle/reference_harness.py) and committed as fixtures under tests/golden/
by tests/golden/make_golden.py.  The block-sparse kernels themselves live in the un-vendored
third-party wheel triton==1.1.0 (requirements.txt:11) which cannot run here; for those the
oracle restates the published algorithm and is anchored on the reference's call sites.
"""
