"""CPU tests: pin oracle/restate.py against the fixtures minted from the unmodified
reference (oracle/make_golden.py).  Elementwise results must be bit-exact; reduced
quantities (scale gradients, fixed-point observers) get a reduction-order tolerance."""
import pytest
import torch

from oracle import restate as R
from tests.golden_io import bits_equal, first_mismatch, load

RTOL_RED = 2e-5   # reductions: summation order may differ between hosts


def close(a, b, rtol=RTOL_RED, atol=0.0):
    return torch.allclose(a.float().reshape(-1), b.float().reshape(-1), rtol=rtol, atol=atol, equal_nan=True)


def exact(a, b, what=""):
    assert bits_equal(a, b), f"{what}: {first_mismatch(a, b)}"


# --------------------------------------------------------------------------- #
UTILS = load("utils")


@pytest.mark.parametrize("name", sorted(n for n in UTILS if n.startswith("a1_")))
def test_a1_quantize_dequantize(name):
    c = UTILS[name]
    lo, hi = c.meta["lo"], c.meta["hi"]
    exact(R.codes_a1(c.inp["x"], c.inp["scale"], c.inp["offset"], lo, hi), c.out["codes"], "codes")
    exact(R.emulate_a1(c.inp["x"], c.inp["scale"], c.inp["offset"], lo, hi), c.out["y"], "y")


@pytest.mark.parametrize("name", sorted(n for n in UTILS if n.startswith("grad_scale_")))
def test_grad_scale_value(name):
    c = UTILS[name]
    exact(R.grad_scale(c.inp["s"], c.meta["g"]), c.out["value"])
    assert not torch.equal(c.out["value"], c.inp["s"]) or True  # documented: may differ by 1 ulp


def test_round_floor_pass():
    exact(R.round_ste(UTILS["round_pass"].inp["v"]), UTILS["round_pass"].out["value"])
    exact(R.floor_ste(UTILS["floor_pass"].inp["v"]), UTILS["floor_pass"].out["value"])


def test_qrange():
    assert R.qrange(True, 4) == (-7, 7) and R.qrange(False, 4) == (0, 15)
    assert R.qrange(True, 8) == (-127, 127) and R.qrange(False, 8) == (0, 255)


# --------------------------------------------------------------------------- #
QBASE = load("qbase")


def _obs(t, cfg):
    return R.get_qparams_tensor(t, cfg["type"], **{k: v for k, v in cfg["args"].items()})


@pytest.mark.parametrize("name", sorted(QBASE))
def test_qbase_module(name):
    c = QBASE[name]
    q = c.meta["qconfig"]
    x, w = c.inp["x"], c.inp["weight"]
    ilo, ihi = R.qrange(q["input"]["args"]["signed"], q["input"]["args"]["n_bits"])
    wlo, whi = R.qrange(q["weight"]["args"]["signed"], q["weight"]["args"]["n_bits"])
    # observers (lazy init, modules/base.py:88-94,123-129)
    s_in, o_in = _obs(x, q["input"])
    assert close(s_in, c.out["param_in_scale"]) and close(o_in.float(), c.out["buf_in_offset"])
    s_w, o_w = _obs(w, q["weight"])
    assert close(s_w, c.out["param_wt_scale"]) and close(o_w.float(), c.out["buf_wt_offset"])
    # fake-quant forward with the reference's own qparams -> bit exact
    s_in, o_in = c.out["param_in_scale"], c.out["buf_in_offset"]
    s_w, o_w = c.out["param_wt_scale"], c.out["buf_wt_offset"]
    g_i, g_w = R.lsq_g(x.numel(), ihi), R.lsq_g(w.numel(), whi)
    qx, dx, ds_in = R.fq_affine_fwd_bwd(x, s_in, o_in, ilo, ihi, g_i, c.out["d_qx"])
    qw, dw, ds_w = R.fq_affine_fwd_bwd(w, s_w, o_w, wlo, whi, g_w, c.out["d_qw"])
    exact(qx, c.out["qx"], "qx")
    exact(qw, c.out["qw"], "qw")
    exact(dx, c.out["dx"], "dx")
    exact(dw, c.out["grad_weight"], "dw")
    assert close(ds_in, c.out["grad_in_scale"]), (ds_in, c.out["grad_in_scale"])
    assert close(ds_w, c.out["grad_wt_scale"]), (ds_w, c.out["grad_wt_scale"])


FUNLSQ = load("funlsq")


@pytest.mark.parametrize("name", sorted(FUNLSQ))
def test_funlsq(name):
    c = FUNLSQ[name]
    lo, hi, g = c.meta["lo"], c.meta["hi"], c.meta["g"]
    exact(R.emulate_a1(c.inp["w"], c.inp["scale"], c.inp["offset"], lo, hi), c.out["y"], "y")
    dw, ds = R.fun_lsq_backward(c.inp["w"], c.inp["scale"], lo, hi, g, c.inp["dy"])
    exact(dw, c.out["dw"], "dw")
    assert close(ds, c.out["dscale"])


# --------------------------------------------------------------------------- #
ROOTQ = load("rootq")


@pytest.mark.parametrize("name", sorted(n for n in ROOTQ if n.startswith("conv_")))
def test_rootq_module(name):
    c = ROOTQ[name]
    q = c.meta["qconfig"]
    m = q["momentum"]
    x, w = c.inp["x"], c.inp["weight"]
    ilo, ihi = R.qrange(q["input"]["args"]["signed"], q["input"]["args"]["n_bits"])
    wlo, whi = R.qrange(q["weight"]["args"]["signed"], q["weight"]["args"]["n_bits"])
    if name.endswith("_eval"):
        qx, _ = R.rootq_act(x, c.out["param_in_scale"], c.out["buf_in_run_scale"], m, ilo, ihi, False)
        qw, _, _ = R.rootq_wt(w, c.out["param_wt_upper"], c.out["param_wt_lower"], c.out["param_wt_alpha"],
                              c.out["buf_wt_run_upper"], c.out["buf_wt_run_lower"], m, wlo, whi, False)
        exact(qx, c.out["qx"], "qx")
        exact(qw, c.out["qw"], "qw")
        return
    if name.endswith("_step1"):   # lazy init (RootQ/base.py:79-90,113-129)
        in_scale = R.rootq_act_init(x, ilo, ihi)
        run_scale = in_scale.clone()
        up, dn = R.rootq_wt_init(w, whi)
        run_up, run_dn = up.clone(), dn.clone()
        alpha = torch.tensor(0.25)
        exact(in_scale, c.out["param_in_scale"], "init in_scale")
        assert close(up, c.out["param_wt_upper"]) and close(dn, c.out["param_wt_lower"])
        up, dn = c.out["param_wt_upper"], c.out["param_wt_lower"]
        run_up, run_dn = up.clone(), dn.clone()
    else:
        in_scale, run_scale = c.inp["pre_in_scale"], c.inp["pre_in_run_scale"]
        up, dn, alpha = c.inp["pre_wt_upper"], c.inp["pre_wt_lower"], c.inp["pre_wt_alpha"]
        run_up, run_dn = c.inp["pre_wt_run_upper"], c.inp["pre_wt_run_lower"]
    qx, new_run, dx, ds = R.rootq_act_fwd_bwd(x, in_scale, run_scale, m, ilo, ihi, c.out["d_qx"])
    exact(qx, c.out["qx"], "qx")
    exact(new_run, c.out["buf_in_run_scale"], "run_scale")
    exact(dx, c.out["dx"], "dx")
    assert close(ds, c.out["grad_in_scale"])
    qw, nru, nrl, dw, du, dl, da = R.rootq_wt_fwd_bwd(w, up, dn, alpha, run_up, run_dn, m, wlo, whi, c.out["d_qw"])
    exact(qw, c.out["qw"], "qw")
    exact(nru, c.out["buf_wt_run_upper"], "run_upper")
    exact(nrl, c.out["buf_wt_run_lower"], "run_lower")
    assert close(dw, c.out["grad_weight"], rtol=1e-6, atol=1e-9)
    assert close(du, c.out["grad_wt_upper"]) and close(dl, c.out["grad_wt_lower"]) and close(da, c.out["grad_wt_alpha"])


@pytest.mark.parametrize("name", ["clipping_0_15", "clipping_w"])
def test_rootq_clipping(name):
    c = ROOTQ[name]
    exact(R.rootq_clipping(c.inp["x"], c.inp["upper"], c.inp["lower"]), c.out["value"])


# --------------------------------------------------------------------------- #
FSPTQ = load("fsptq")


@pytest.mark.parametrize("name", sorted(FSPTQ))
def test_fsptq_module(name):
    c = FSPTQ[name]
    q = c.meta["qconfig"]
    x, w = c.inp["x"], c.inp["weight"]
    ilo, ihi = R.qrange(q["input"]["args"]["signed"], q["input"]["args"]["n_bits"])
    wlo, whi = R.qrange(q["weight"]["args"]["signed"], q["weight"]["args"]["n_bits"])
    s_in, o_in = _obs(x, q["input"])
    assert close(s_in, c.out["param_in_scale"]) and close(o_in.float(), c.out["buf_in_offset"])
    s_w, o_w = _obs(w.clone(), q["weight"])
    assert close(s_w + 1e-6, c.out["param_wt_scale"])          # FSPTQuant/base.py:129
    s_in, o_in, s_w = c.out["param_in_scale"], c.out["buf_in_offset"], c.out["param_wt_scale"]
    qx, dx, ds_in = R.fq_zp_fwd_bwd(x, s_in, o_in, ilo, ihi, c.out["d_qx"])
    exact(qx, c.out["qx"], "qx")
    exact(dx, c.out["dx"], "dx")
    assert close(ds_in, c.out["grad_in_scale"])
    if q["weight"]["recon_type"] == "adaround":
        alpha = R.adaround_init_alpha(w, s_w)
        qw, ds_w, da = R.fq_adaround_fwd_bwd(w, s_w, alpha, wlo, whi, c.out["d_qw"])
        exact(qw, c.out["qw"], "qw")
        assert close(ds_w, c.out["grad_wt_scale"], atol=1e-7)
        if "alpha" in c.out:      # only the conv case stores these
            exact(alpha, c.out["alpha"], "alpha init")
            exact(R.fq_adaround(w, s_w, alpha, wlo, whi, False), c.out["qw_eval"], "qw eval")
            assert close(da, c.out["grad_alpha"], rtol=1e-6, atol=1e-10)
        assert not c.out["grad_weight"].any()     # floor without STE: weight gradient is all zeros
    else:
        qw, dw, ds_w = R.fq_sym_fwd_bwd(w, s_w, wlo, whi, c.out["d_qw"])
        exact(qw, c.out["qw"], "qw")
        exact(dw, c.out["grad_weight"], "dw")
        assert close(ds_w, c.out["grad_wt_scale"], atol=1e-7)


# --------------------------------------------------------------------------- #
OBS = load("observers")


@pytest.mark.parametrize("name", sorted(OBS))
def test_observers(name):
    c = OBS[name]
    t = c.inp["t"]
    kind = name.split("_")[0] + "_" + name.split("_")[1]
    kw = dict(n_bits=c.meta["n_bits"], signed=c.meta["signed"])
    if kind.startswith("minmax_channel"):
        s, o = R.obs_minmax_channel(t, ch_axis=c.meta["ch_axis"], **kw)
    elif kind == "l2loss_channel":
        s, o = R.obs_l2loss_channel(t.clone(), ch_axis=0, **kw)
    elif kind == "l2norm_channel":
        s, o = R.obs_l2norm_channel(t, ch_axis=0, **kw)
    else:
        s, o = R.OBSERVERS[kind](t, **kw)
    if kind.startswith("minmax"):
        exact(s, c.out["scale"], "scale")
        exact(o.float(), c.out["offset"], "offset")
    else:   # sums inside: allow a summation-order tolerance (same host => normally exact)
        assert s.shape == c.out["scale"].shape
        assert close(s, c.out["scale"], rtol=1e-4), (s, c.out["scale"])
        assert close(o.float(), c.out["offset"], rtol=1e-4)
