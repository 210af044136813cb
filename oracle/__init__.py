"""Test infrastructure only — NOT part of the product path.

`oracle/` holds CPU restatements of the reference FieldConv hot path
(twmitchel/FieldConv, nn/field_conv.py:104-137 and helpers) used as the
checker by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs.  Nothing under `fieldconv_b200/`
may import it.

Parity pinning: the reference ships no tests or golden vectors
(SURVEY.md §4), so the restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF, executed unmodified in the build container through
`oracle/ref_loader.py`; the resulting vectors are committed under
`tests/golden/` together with the generating script
(`oracle/make_golden.py`).
"""
