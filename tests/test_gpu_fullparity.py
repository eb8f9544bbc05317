"""Deterministic-mode parity at BASELINE.json's own sizes: CUDA path through
the C-ABI vs the oracle's de-duplicated corrected mode - allocations, ancestors, selected particle,
cluster sizes and resampling count bit-exact; log-probs / log-weights within 1e-5 relative
(north_star); and the pool engine performs exactly the reference's calc_logprob calls."""
import pytest

import fullsize

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(fullsize.FULL_CASES))
def test_full_size_parity(name):
    import pmdi_b200.capi as capi
    cfg = fullsize.make(name)
    hy = cfg["hy"]
    with capi.Context(cfg["data"], cfg["types"], cfg["N"], cfg["P"]) as ctx:
        def run(s, order, it, lw0, debug):
            return ctx.sweep(s, order, cfg["n1"], hy["Pi"], hy["phi"], seed=77, it=it, logweight_init=lw0, debug=debug)
        out = fullsize.compare(cfg, run)
    print(name, out)
    fullsize.assert_parity(out)
    if name in ("cfg2_multiomics_fresh", "cfg3_tcga"):
        assert out["n_resamples"][0] > 0  # the resampling path ran at full size
