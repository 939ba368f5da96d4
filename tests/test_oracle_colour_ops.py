"""The C oracle's colour operators (f32) against an independent float64 numpy restatement of the reference's shaders:
linear.frag:12-17 (3x3), oklab.frag:34-64 (M1 / M2 as quoted in SURVEY.md appendix A.5, cube root, in-shader inverses) and
srlab2.frag:36-120 (CAT02 / HPE, whitepoint ignored on encode and applied on decode, the 216/24389 non-linearity).
The perceptual goldens pin these only to a 16 x 16 block hash; this pins every constant.  Tolerance: 1e-5 relative +
2e-6 absolute (f32 evaluation of 3-4 chained 3x3 products)."""
import numpy as np

from oracle import oracle as O

M1 = np.array([[0.8189330101, 0.3618667424, -0.1288597137], [0.0329845436, 0.9293118715, 0.0361456387],
               [0.0482003018, 0.2643662691, 0.6338517070]])
M2 = np.array([[0.2104542553, 0.7936177850, -0.0040720468], [1.9779984951, -2.4285922050, 0.4505937099],
               [0.0259040371, 0.7827717662, -0.8086757660]])
CAT02 = np.array([[0.7328, 0.4296, -0.1624], [-0.7036, 1.6975, 0.0061], [0.0030, 0.0136, 0.9834]])
HPE = np.array([[0.38971, 0.68898, -0.07868], [-0.22981, 1.18340, 0.04641], [0.0, 0.0, 1.0]])


def close(got, exp, rel=1e-5, ab=2e-6):
    return np.all(np.abs(got.astype(np.float64) - exp) <= rel * np.abs(exp) + ab)


def rgba(n, seed):
    return np.random.default_rng(seed).uniform(0, 1, (1, n, 4)).astype(np.float32)


def test_linear_matrix():
    tex = rgba(5000, 1)
    M = np.random.default_rng(2).uniform(-1, 1, (3, 3))
    got = O.linear(tex, M)[0]
    assert close(got[:, :3], tex[0, :, :3].astype(np.float64) @ M.astype(np.float32).astype(np.float64).T)
    assert np.array_equal(got[:, 3], tex[0, :, 3])  # alpha passes through


def test_oklab_encode_decode():
    T = O.to_xyz("bt709", "D65").astype(np.float64).reshape(3, 3)
    tex = rgba(5000, 3)
    rgb = tex[0, :, :3].astype(np.float64)
    lab = np.cbrt((rgb @ T.T) @ M1.T) @ M2.T
    got = O.oklab_encode(tex, T)[0]
    assert close(got[:, :3], lab) and np.array_equal(got[:, 3], tex[0, :, 3])
    # white maps to L = 1, a = b = 0 (Ottosson's normalisation) -- a check of the constants themselves
    white = O.oklab_encode(np.ones((1, 1, 4), np.float32), T)[0, 0]
    assert abs(white[0] - 1) < 1e-4 and abs(white[1]) < 1e-4 and abs(white[2]) < 1e-4
    # decode: (M2^-1 Lab)^3, M1^-1, T^-1, clamp; fed with the oracle's own f32 Lab values
    Ti = np.linalg.inv(T)
    lab32 = np.concatenate([got[:, :3], tex[0, :, 3:]], 1)[None]
    back = O.oklab_decode(lab32, Ti)[0]
    exp = np.clip((((got[:, :3].astype(np.float64) @ np.linalg.inv(M2).T) ** 3) @ np.linalg.inv(M1).T) @ Ti.T, 0, 1)
    assert close(back[:, :3], exp, ab=1e-5) and np.array_equal(back[:, 3], tex[0, :, 3])
    assert np.all(np.abs(back[:, :3] - tex[0, :, :3]) < 2e-5)  # round trip inside the gamut
    # outside the gamut the decode clamps to [0, 1]
    wild = O.oklab_decode(np.array([[[0.9, 0.4, -0.4, 0.5]]], np.float32), Ti)[0, 0]
    assert wild[:3].min() >= 0 and wild[:3].max() <= 1 and wild[3] == 0.5


def nonlin(v):
    return np.where(np.abs(v) < 216 / 24389, v * 24389 / 2700, 1.16 * np.cbrt(v) - 0.16)


def nonlin_inv(v):
    return np.where(np.abs(v) < 0.08, v * 2700 / 24389, ((v + 0.16) / 1.16) ** 3)


def test_srlab2_encode_decode():
    T = O.to_xyz("bt709", "D65").astype(np.float64).reshape(3, 3)
    tex = rgba(5000, 4)
    tex[0, :200, :3] *= 0.01  # the linear toe of the non-linearity
    rgb = tex[0, :, :3].astype(np.float64)
    lms = (((rgb @ T.T) @ CAT02.T) @ np.linalg.inv(CAT02).T) @ HPE.T  # wp_rgb = 1: the whitepoint is ignored on encode
    e = nonlin(lms) @ np.linalg.inv(HPE).T
    lab = np.stack([e[:, 1], (e[:, 0] - e[:, 1]) * 5 / 1.16, (e[:, 2] - e[:, 1]) * 2 / 1.16], 1)
    got = O.srlab2_encode(tex, T)[0]
    assert close(got[:, :3], lab, ab=1e-5) and np.array_equal(got[:, 3], tex[0, :, 3])

    wp = np.array(O.WHITEPOINTS["D65"])
    Ti = np.linalg.inv(T)
    g = got[:, :3].astype(np.float64)
    xe = np.stack([g[:, 1] * 1.16 / 5 + g[:, 0], g[:, 0], g[:, 2] * 1.16 / 2 + g[:, 0]], 1)
    rgb_w = (nonlin_inv(xe @ HPE.T) @ np.linalg.inv(HPE).T) @ CAT02.T
    exp = np.clip(((rgb_w * (CAT02 @ wp)) @ np.linalg.inv(CAT02).T) @ Ti.T, 0, 1)  # the whitepoint is applied on decode
    lab32 = np.concatenate([got[:, :3], tex[0, :, 3:]], 1)[None]
    back = O.srlab2_decode(lab32, Ti, wp)[0]
    assert close(back[:, :3], exp, ab=2e-5) and np.array_equal(back[:, 3], tex[0, :, 3])


def test_oklab_published_vectors():
    """Ottosson's table of example XYZ -> Oklab pairs ("A perceptual color space for image processing", 2020), given to three
    decimals; with T = identity the operator takes XYZ directly."""
    pairs = [((0.950, 1.000, 1.089), (1.000, 0.000, 0.000)), ((1.000, 0.000, 0.000), (0.450, 1.236, -0.019)),
             ((0.000, 1.000, 0.000), (0.922, -0.671, 0.263)), ((0.000, 0.000, 1.000), (0.153, -1.415, -0.449))]
    tex = np.array([[list(x) + [1.0] for x, _ in pairs]], np.float32)
    got = O.oklab_encode(tex, np.eye(3))[0, :, :3]
    assert np.abs(got - np.array([l for _, l in pairs])).max() < 1.5e-3
