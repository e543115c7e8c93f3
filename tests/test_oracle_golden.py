"""The C oracle (oracle/ta_oracle.c) against outputs of the unmodified reference (tests/golden/*.npz,
made by tests/golden/make_golden.py).  Bit-exact wherever the reference's op order is fully determined
by its Python expressions; explicit tolerances (stated per test) where ATen's internal summation /
FMA-contraction order is not (bilinear blend, depthwise conv, L2 norms, per-sample mean)."""
import numpy as np
import pytest

import oracle
from conftest import bits_equal, load_golden, n_diff_bits, ulp_diff


@pytest.fixture(scope="module")
def H():
    return load_golden("hooks")


def test_abs_mean_vs_torch(H):
    mine = oracle.abs_mean_per_sample(H["g"])
    # torch's CPU mean uses its own vectorised fp32 summation order; the exact (fp64) mean may differ by 1 ulp
    assert ulp_diff(mine, H["scale"]).max() <= 1


@pytest.mark.parametrize("key,decay,first", [("mom_first", 1.0, True), ("mom_d1", 1.0, False),
                                             ("mom_d07", 0.7, False), ("mom_d0", 0.0, False)])
def test_momentum_bit_exact(H, key, decay, first):
    out = oracle.momentum(H["g"], None if first else H["m"], H["scale"], decay)
    assert bits_equal(out, H[key]), n_diff_bits(out, H[key])


def test_momentum_nan_sample(H):
    scale = np.array(H["scale"]); scale[1] = 0.0
    out = oracle.momentum(H["gz"], H["m"], scale, 1.0)
    assert bits_equal(out, H["mom_nan"])
    assert np.isnan(out[1]).all()


def test_update_linf_bit_exact(H):
    eps, alpha = float(H["eps"]), float(H["alpha"])
    for key, kw in [("upd_linf", dict(alpha=alpha)), ("upd_linf_neg", dict(alpha=-alpha)),
                    ("upd_linf_tensor", dict(alpha=0.0, alpha_t=H["alpha_t"]))]:
        out = oracle.update_linf(H["delta"], H["data"], H["mom_d1"], eps=eps, **kw)
        assert bits_equal(out, H[key]), (key, n_diff_bits(out, H[key]))
    out = oracle.update_linf(H["delta"], H["data"], H["mom_nan"], alpha=alpha, eps=eps)
    assert bits_equal(out, H["upd_nan"])


def test_fused_update_matches_reference_pair(H):
    eps, alpha = float(H["eps"]), float(H["alpha"])
    m, d, x = oracle.fused_update_linf(H["g"], H["m"], H["delta"], H["data"], H["scale"], 1.0, alpha, eps)
    assert bits_equal(m, H["mom_d1"]) and bits_equal(d, H["upd_linf"])
    assert bits_equal(x, (H["data"] + d).astype(np.float32))
    m, d, _ = oracle.fused_update_linf(H["g"], None, H["delta"], H["data"], H["scale"], 1.0, alpha, eps)
    assert bits_equal(m, H["mom_first"])


def test_update_l2(H):
    eps = float(H["eps"])
    out = oracle.update_l2(H["delta"] * np.float32(0.01), H["data"], H["g_l2"], 0.01, eps)
    # norms are fp64-accumulated here, fp32 in torch: tolerance 2e-7 abs on values of O(0.06)
    np.testing.assert_allclose(out, H["upd_l2_small"], rtol=0, atol=2e-7)
    out = oracle.update_l2(H["delta"], H["data"], H["g_l2"], 2.0, eps)
    np.testing.assert_allclose(out, H["upd_l2_big"], rtol=0, atol=2e-7)


def test_init_delta(H):
    out = oracle.clamp_box(H["init_noise"], H["data"])
    assert bits_equal(out, H["init_linf"])
    out = oracle.init_l2_scale(H["init_l2_normal"], H["init_l2_r"], H["data"], float(H["eps"]))
    np.testing.assert_allclose(out, H["init_l2"], rtol=0, atol=1e-8)


def test_stage_add_and_ni(H):
    assert bits_equal(oracle.stage_add(H["data"], H["delta"]), H["x_adv"])
    out = oracle.stage_add(H["data"], H["delta"], H["m"], float(H["ni_coef"]))
    assert bits_equal(out, H["ni_x"])


def test_normalize():
    M = load_golden("misc")
    assert bits_equal(oracle.normalize_fwd(M["norm_x"], M["norm_mean"], M["norm_std"]), M["norm_y"])
    assert bits_equal(oracle.normalize_bwd(M["norm_gout"], M["norm_std"]), M["norm_gin"])


def test_quantize_u8():
    M = load_golden("misc")
    out = oracle.quantize_u8(M["q_data"], M["q_delta"], to_nhwc=True)
    assert np.array_equal(out, M["q_u8"])
    assert list(out[0, 0, :4, 0]) == [254, 127, 254, 255]


def test_sim():
    G = load_golden("sim_admix_emi")
    S = int(G["sim_S"])
    assert bits_equal(oracle.sim_fwd(G["sim_x"], S), G["sim_y"])
    assert bits_equal(oracle.sim_bwd(G["sim_gout"], S), G["sim_gin"])


def test_admix():
    G = load_golden("sim_admix_emi")
    S, A = int(G["admix_S"]), int(G["admix_A"])
    out = oracle.admix_fwd(G["sim_x"], G["admix_perm"], float(G["admix_strength"]), S)
    assert bits_equal(out, G["admix_y"])
    gin = oracle.admix_bwd(G["admix_gout"], S, A)
    assert bits_equal(gin, G["admix_gin"]), n_diff_bits(gin, G["admix_gin"])


def test_emi():
    G = load_golden("sim_admix_emi")
    K = G["emi_coef"].size
    assert bits_equal(oracle.lin_sample_fwd(G["sim_x"], G["emi_gbar"], G["emi_coef"]), G["emi_y"])
    assert bits_equal(oracle.lin_sample_fwd(G["sim_x"], None, G["emi_coef"]), G["emi_y0"])
    gin = oracle.lin_sample_bwd(G["emi_gout"], K)
    assert bits_equal(gin, G["emi_gin"]), n_diff_bits(gin, G["emi_gin"])


def test_vmi():
    V = load_golden("vmi")
    N = int(V["N"])
    acc = None
    for k in range(N):
        xn = oracle.neighbor_stage(V["data"], V["delta"], V["noises"][k])
        assert bits_equal(xn, V["x_near"][k])
        acc = oracle.accumulate(acc, V["grads"][k], first=(k == 0))
    var = oracle.variance_finalize(acc, V["cur"], N)
    assert bits_equal(var, V["variance"])
    assert bits_equal(oracle.add(V["cur"], var), V["g_plus_v"])


def _dim_cases():
    D = load_golden("dim")
    return D, sorted({k.rsplit("_", 1)[0] for k in D.files})


def test_dim_forward():
    D, cases = _dim_cases()
    assert len(cases) >= 7
    for c in cases:
        rnd, R, top, left, _ = [int(v) for v in D[c + "_params"]]
        out = oracle.dim_fwd(D[c + "_x"], rnd, R, top, left)
        # ATen contracts the blend into FMAs at its compiler's discretion: 1-2 ulp, values in [0,1]
        np.testing.assert_allclose(out, D[c + "_y"], rtol=0, atol=3e-7, err_msg=c)


def test_dim_backward():
    D, cases = _dim_cases()
    for c in cases:
        rnd, R, top, left, _ = [int(v) for v in D[c + "_params"]]
        gin = oracle.dim_bwd(D[c + "_gout"], rnd, R, top, left)
        np.testing.assert_allclose(gin, D[c + "_gin"], rtol=0, atol=2e-6, err_msg=c)


def _tim_kernels():
    T = load_golden("tim")
    return T, sorted(k[:-7] for k in T.files if k.endswith("_kernel"))


def test_tim_conv_2d():
    T, kernels = _tim_kernels()
    for key in kernels:
        k = T[key + "_kernel"]
        for tag in "abc":
            if key + "_" + tag + "_in" not in T.files:
                continue
            out = oracle.dwconv2d(T[key + "_" + tag + "_in"], k)
            # 225-term fp32 sums of N(0,1)*weights(sum 1): both sides round differently; 1e-6 abs
            np.testing.assert_allclose(out, T[key + "_" + tag + "_out"], rtol=0, atol=1e-6, err_msg=key + tag)


def test_tim_conv_separable_matches_2d():
    import transferattack_b200.input_transformation.tim as tim
    T, kernels = _tim_kernels()
    for key in kernels:
        kt, ks = key.rstrip("0123456789"), int(key[len(key.rstrip("0123456789")):])
        k2d, kcol, krow = tim.make_kernel(kt, ks)
        assert bits_equal(k2d, T[key + "_kernel"]), key           # generate_kernel parity (tim.py:42-66)
        x = T[key + "_a_in"]
        out = oracle.dwconv2d_sep(x, np.stack([kcol] * 3), np.stack([krow] * 3))
        np.testing.assert_allclose(out, T[key + "_a_out"], rtol=0, atol=1e-6, err_msg=key)


def test_philox_restatement_known_answers():
    """oracle/philox.py: Philox4x32-10 against the Random123 known-answer vectors (Salmon et al., SC'11, kat_vectors), and
    the execution policy / element mapping of the uniform fill on a B200-shaped device (148 SMs x 2048 threads)."""
    from oracle import philox as P

    def kat(c, k):
        o = P.philox4x32_10(*[np.array([v], np.uint64) for v in c], k[0], k[1])
        return tuple(int(v[0]) for v in o)
    assert kat((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert kat((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert kat((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    assert P.torch_uniform_policy(64 * 3 * 224 * 224, 148, 2048) == (303104, 32)
    assert P.torch_uniform_policy(1000, 148, 2048) == (1024, 4)
    v = P.torch_uniform(5000, seed=1234, offset=8, frm=-0.094, to=0.094, T=1024)
    assert v.dtype == np.float32 and v.min() >= np.float32(-0.094) and v.max() < np.float32(0.094)
    assert abs(float(v.mean())) < 0.01 and len(np.unique(v)) > 4900
