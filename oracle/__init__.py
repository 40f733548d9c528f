"""CPU oracle for the BOA hot path (TEST INFRASTRUCTURE ONLY).

A numpy / torch-fp32 restatement of the reference algorithm for the path
`3D U-Net sliding-window inference -> Gaussian aggregation -> argmax label map -> HU tissue rules ->
per-label / per-slice reductions`.  Every function cites the reference file:line it follows
(paths relative to /root/reference/body_organ_analysis).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
this package, and only as the checker or the timed CPU baseline - never the product path.

Parity pinning (see DESIGN.md "Oracle"):
  * sliding-window geometry, Gaussian map, CT normalisation, class maps, tissue rules, slice/label statistics and
    both JSON builders are pinned against the reference's OWN functions, imported from /root/reference with small
    stubs by `tests/golden/make_golden.py`; the outputs are committed under `tests/golden/`.
  * the network (un-vendored PyPI `dynamic-network-architectures==0.4.3`) is restated from its published structure
    on `torch.nn.functional`; the reference holds no golden logits / label maps for it, so that part
    is "parity unpinned" - the restatement is anchored on the reference's call sites only.
"""
