"""CPU oracle of the Multi-HMR hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch fp32 restatement of `Model.forward(x, K)` of naver/multi-hmr (reference
model.py:205-349) and of the un-vendored third-party arithmetic it calls (facebookresearch/dinov2
hub ViT, `smplx` body model + LBS, `roma` rotation conversions).  Each function cites the reference
file:line (or, for the third-party code that is absent from /root/reference, the published algorithm)
that it follows.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import this package, and only as the checker or the timed CPU baseline — never as the product path.

Parity pinning: the reference ships no tests or golden vectors.  The restatement is
pinned against the reference's OWN Python, imported unmodified from /root/reference in the build
container by `oracle/make_golden.py` (which shims only the three missing third-party packages with the
restatements in this package) — see tests/golden/ and tests/test_oracle_golden.py.  The third-party
restatements themselves (dinov2 / smplx / roma) have no upstream source in this environment to be
checked against: for those three the status is "parity unpinned" (DESIGN.md §Oracle).
"""
