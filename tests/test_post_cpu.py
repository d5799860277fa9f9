'''
CPU tests of the post-processing that consumes the hot path's outputs -- the windowed r_eff methods and the generation time
(reference sim.py:888-1025) -- as functions of plain arrays, against the values the unmodified reference computed for the
golden scenarios (oracle/gen_golden.py stores r_eff/infectious, r_eff/outcome and gen_time next to the People arrays and the
infection log they are computed from).
'''
import numpy as np
import pytest

import scenarios

FULL = ['hybrid3k', 'random2k_nowaning', 'variants4k', 'dynamic2k', 'dynpars3k', 'clip3k', 'rescale3k', 'fracsus2k', 'sequence3k',
        'testnum3k', 'subtarget3k', 'capacity3k']


@pytest.mark.parametrize('name', FULL)
def test_r_eff_methods_and_gen_time(name, golden):
    from covasim_b200.sim import r_eff_windowed, gen_time
    g = golden(name)
    npts = scenarios.SCENARIOS[name]['pars'].get('n_days', 60) + 1
    for method in ('infectious', 'outcome'):
        got = r_eff_windowed(method, g['people/date_infectious'], g['people/date_recovered'], g['people/date_dead'], g['log/source'], npts)
        np.testing.assert_allclose(got, g[f'r_eff/{method}'], rtol=1e-12, atol=0, equal_nan=True, err_msg=f'{name} {method}')
        assert np.isfinite(got).sum() > npts // 2
    gt = gen_time(g['people/date_exposed'], g['people/date_symptomatic'], g['log/source'], g['log/target'])
    np.testing.assert_allclose([gt['true'], gt['true_std'], gt['clinical'], gt['clinical_std']], g['gen_time'], rtol=1e-12)
    assert np.isnan(gt['true']) or 2 < gt['true'] < 15          # NaN when sources were made naive again (rescaling), as in the reference
