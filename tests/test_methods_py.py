"""Python mirror of the method API: host-side formulas agree with the oracle's restatement of the reference."""
import numpy as np
import pytest

from oracle import oracle as o


def test_get_err_formula_and_true_price_line():
    from nmch_b200 import methods as M
    obj = M.NMCH.__new__(M.NMCH_FE_K3_MM)       # host fields only: no engine, no GPU
    obj.state_numbers, obj.strike_price, obj.price_squared = 512 * 512, 0.120281939, 0.045731643
    obj.S_0, obj.K, obj.r, obj.sigma = 1.0, 1.0, 0.0, 0.3
    assert abs(obj.get_err() - o.get_err(512 * 512, 0.120281939, 0.045731643)) < 1e-9
    assert abs(obj.get_err() - 0.000819) < 2e-6
    assert abs(obj.true_price_line() - o.lib().orc_print_true_price(1.0, 1.0, 0.0, 0.3)) < 1e-7


def test_unsupported_tag_is_loud():
    from nmch_b200 import methods as M
    with pytest.raises(ValueError):
        M.NMCH_FE_K3_MM(512, 8, 1.0, 1.0, 0.1, 0.0, 0.5, -0.7, 0.1, 0.3, 100, "curandStateSobol32_t")


@pytest.mark.gpu
def test_python_method_objects_match_engine_and_text_format():
    from nmch_b200 import engine as E
    from nmch_b200 import methods as M
    for cls, method in ((M.NMCH_FE_K3_MM, E.METHOD_FE), (M.NMCH_EM_K3_MM, E.METHOD_EM)):
        m = cls(512, 16, 1.0, 1.0, 0.1, 0.0, 0.5, -0.7, 0.1, 0.3, 100, M.XORWOW)
        m.init(1234)
        m.compute()
        m.set_k(1.5)
        m.compute()
        with E.Engine(NTPB=512, NB=16, N=100, method=method, rng=E.RNG_XORWOW_COMPAT) as e:
            e.init(1234)
            e.compute()
            e.set_params(1.5, 0.1, 0.3)
            want = e.compute()
        assert m.last_moments.sum_payoff == want.sum_payoff
        assert m.get_strike_price() == float(np.float32(want.mean))
        text = m.stats_text().splitlines()
        assert text[0] == "Base parameters:" and text[7] == "k       = 1.500000"
        assert text[12] in ("METHOD: FORWARD-EULER", "METHOD: EXACT-METHOD")
        assert abs(m.get_err() - o.get_err(512 * 16, m.get_strike_price(), m.get_price_squared())) < 1e-9
        m.finalize()
        m.finalize()


@pytest.mark.gpu
def test_python_opt_in_fe_streams_select_the_engine_modes():
    """dense= (Philox tag) and fast= (XORWOW tag) pick the opt-in FE streams; EM classes ignore them, as in C++."""
    from nmch_b200 import engine as E
    from nmch_b200 import methods as M
    args = (512, 16, 1.0, 1.0, 0.1, 0.0, 0.5, -0.7, 0.1, 0.3, 100)
    for tag, kw, mode in ((M.PHILOX, dict(dense=True), E.RNG_PHILOX_DENSE), (M.XORWOW, dict(fast=True), E.RNG_XORWOW_FAST)):
        m = M.NMCH_FE_K3_MM(*args, tag, **kw)
        m.init(1234)
        m.compute()
        with E.Engine(NTPB=512, NB=16, N=100, rng=mode) as e:
            e.init(1234)
            want = e.compute()
        assert m.last_moments.sum_payoff == want.sum_payoff
        m.finalize()
    m = M.NMCH_EM_K3_MM(*args, M.XORWOW, fast=True)            # FE-only option: the EM class stays draw-compatible
    m.init(1234)
    m.compute()
    with E.Engine(NTPB=512, NB=16, N=100, method=E.METHOD_EM, rng=E.RNG_XORWOW_COMPAT) as e:
        e.init(1234)
        want = e.compute()
    assert m.last_moments.sum_payoff == want.sum_payoff
    m.finalize()
