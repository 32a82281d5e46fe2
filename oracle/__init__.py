"""CPU oracle for the SPI inversion hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package.  `spi_b200/` never does: the product path has no CPU fallback.

What it is: a plain-PyTorch (CPU, fp32) functional restatement of the reference algorithms on the
path named by BASELINE.json (`SURVEY.md` §8a rows a1-a26).  Every function cites the reference
file:line it follows.  Arithmetic that the reference itself delegates to torch / torchvision
(`F.conv2d`, `F.grid_sample`, `roi_align`, `torch.sort`, `torch.searchsorted`) is delegated to the
same library calls here, as SURVEY.md §8c prescribes ("oracle = the versions in this container").

Pinning: the reference ships no tests and no golden vectors, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by `oracle/make_golden.py`
(which imports `/root/reference` read-only through `oracle/ref_shim.py`) and committed under
`tests/golden/`.  `tests/test_oracle_golden.py` re-checks the oracle against those fixtures on CPU.
"""
