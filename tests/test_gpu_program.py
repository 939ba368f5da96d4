"""GPU tests at the program level: the pipelines of the reference's integration tests
(lib/zosimos/tests/blend.rs, knobs.rs, direct.rs) written against the mirror API
(CommandBuffer -> Linker.compile -> Program.lower_to -> Executable.launch -> step -> Retire),
checked against the reference's golden hashes and, byte for byte, against the oracle; plus the
planner's fusion (fused == unfused in ZOS_FUSE_EXACT mode)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import refpipes as R
from tests.gpu_common import oracle_desc, oracle_image

pytestmark = pytest.mark.gpu

import zosimos_b200 as Z  # noqa: E402
from zosimos_b200 import _ffi  # noqa: E402
from zosimos_b200.buffer import Color, Descriptor, SampleBits, SampleParts, Texel, Transfer  # noqa: E402
from zosimos_b200.command import (Affine, AffineSample, Bilinear, Blend, ChromaticAdaptationMethod, CommandBuffer, Derivative,  # noqa: E402
                                  DerivativeMethod, Linker, Palette, Rectangle, RegisterKnob, ResizeMode)
from zosimos_b200.program import Capabilities, Pool, StartError, StepError  # noqa: E402


@pytest.fixture(scope="module")
def pool():
    p = Pool()
    p.request_device(0)
    yield p
    for c in p.iter_devices():
        c.close()


def run_once_with_output(commands, pool, binds, out_reg, fuse_mode=_ffi.FUSE_EXACT, knobs=()):
    """tests/util.rs:62-117."""
    plan = Linker.from_included().compile(commands)
    caps = Capabilities.from_device(next(pool.iter_devices()), fuse_mode)
    executable = plan.lower_to(caps)
    return run_executable_with_output(executable, pool, binds, out_reg, knobs)


def run_executable_with_output(executable, pool, binds, out_reg, knobs=()):
    env = executable.from_pool(pool)
    for knob, data in knobs:
        env.knob(knob, data)
    for reg, key in binds:
        env.bind(reg, key)
    env.recover_buffers()
    execution = executable.launch(env)
    pool.clear_cache()
    n = 0
    while execution.is_running():
        execution.step().block_on()
        n += 1
    retire = execution.retire_gracefully(pool)
    img = retire.output(out_reg)
    retire.retire_buffers()
    retire.finish()
    return img, n


def hashes():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_hashes.json")) as f:
        return json.load(f)


def rgba(img):
    d = img.descriptor()
    return img.as_bytes().reshape(d.layout.height, d.layout.width, 4)


@pytest.fixture(scope="module")
def images(pool, fixtures):
    bg = pool.insert_srgb(fixtures["background"])
    fg = pool.insert_srgb(fixtures["foreground"])
    return bg, fg


def test_run_blending(pool, images, fixtures):  # tests/blend.rs:81-116
    bg, fg = images
    c = CommandBuffer()
    background, foreground = c.input(bg.descriptor()), c.input(fg.descriptor())
    placement = Rectangle(0, 0, fg.layout().width, fg.layout().height)
    result = c.inscribe(background, placement, foreground)
    output, _ = c.output(result)
    img, n = run_once_with_output(c, pool, [(background, bg.key()), (foreground, fg.key())], output)
    assert n == 1  # the reference needs 2 draws + staging copies; here: one kernel
    assert O.blockhash256(rgba(img)) in hashes()["composed"]
    exp = O.inscribe(oracle_image(bg.descriptor(), fixtures["background"]), (0, 0, 157, 151), oracle_image(fg.descriptor(), fixtures["foreground"]))
    assert np.array_equal(img.as_bytes(), exp.data.reshape(-1))


def test_run_affine(pool, images, fixtures):  # tests/blend.rs:118-166
    bg, fg = images
    fw, fh, W, H = fg.layout().width, fg.layout().height, bg.layout().width, bg.layout().height
    affine = Affine.new(AffineSample.Nearest).shift(-float(fw // 2), -float(fh // 2)).rotate(np.float32(np.pi) / np.float32(4)).shift(float(W // 2), float(H // 2))
    c = CommandBuffer()
    background, foreground = c.input(bg.descriptor()), c.input(fg.descriptor())
    output, _ = c.output(c.affine(background, affine, foreground))
    img, _ = run_once_with_output(c, pool, [(background, bg.key()), (foreground, fg.key())], output)
    assert O.blockhash256(rgba(img)) in hashes()["affine"]
    exp = O.affine(oracle_image(bg.descriptor(), fixtures["background"]), np.array(affine.transformation, np.float32).reshape(3, 3),
                   oracle_image(fg.descriptor(), fixtures["foreground"]))
    assert np.array_equal(img.as_bytes(), exp.data.reshape(-1))


def test_run_adaptation(pool, images, fixtures):  # tests/blend.rs:168-195
    bg, _ = images
    c = CommandBuffer()
    background = c.input(bg.descriptor())
    output, fmt = c.output(c.chromatic_adaptation(background, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D50))
    assert fmt.color.whitepoint == Z.Whitepoint.D50
    img, _ = run_once_with_output(c, pool, [(background, bg.key())], output)
    assert O.blockhash256(rgba(img)) in hashes()["adapted"]
    exp = O.chromatic_adaptation(oracle_image(bg.descriptor(), fixtures["background"]), "vonkries", "D50")
    assert np.array_equal(img.as_bytes(), exp.data.reshape(-1))


def test_run_conversion(pool, images, fixtures):  # tests/blend.rs:197-224
    bg, _ = images
    bt = pool.allocate_like(bg.key())
    bt.set_color(Color.BT709_RGB)
    c = CommandBuffer()
    inp = c.input(bt.descriptor())
    output, _ = c.output(c.color_convert(inp, bg.descriptor().color, bg.descriptor().texel))
    img, _ = run_once_with_output(c, pool, [(inp, bt.key())], output)
    assert O.blockhash256(rgba(img)) in hashes()["convert_bt709"]
    exp = O.color_convert(oracle_image(bt.descriptor(), fixtures["background"]), O.SRGB, O.RGBA8)
    d = np.abs(img.as_bytes().astype(int) - exp.data.reshape(-1).astype(int))
    assert d.max() <= 1 and np.mean(d == 0) > 0.99


@pytest.mark.parametrize("space", ["oklab", "srlab2"])
@pytest.mark.parametrize("fuse", [_ffi.FUSE_EXACT, _ffi.FUSE_NONE])
def test_run_oklab_srlab2(pool, space, fuse):  # tests/blend.rs:449-580
    c = CommandBuffer()
    color_descriptor = Descriptor.with_srgb_image("rgba8", 400, 400)
    distribution_layout = color_descriptor.with_color(Color.Scalars(Transfer.Linear))
    model = Color.Oklab if space == "oklab" else Color.SrLab2(Z.Whitepoint.D65)
    lab_texel = Descriptor(distribution_layout.layout, model, Texel(Z.Block.Pixel, SampleBits.UInt8x4, SampleParts.LchA))
    g = R.LCH_GRID
    grid = c.bilinear(distribution_layout, Bilinear(g[0], g[1], g[2], g[3], g[4], g[5]))
    lch = c.transmute(grid, lab_texel)
    converted = c.color_convert(lch, color_descriptor.color, color_descriptor.texel)
    output, _ = c.output(converted)
    img, n = run_once_with_output(c, pool, [], output, fuse)
    assert n == 3
    assert O.blockhash256(rgba(img)) in hashes()[space]


def test_run_solid(pool):  # tests/blend.rs:623-644
    c = CommandBuffer()
    output, _ = c.output(c.solid_rgba(Descriptor.with_srgb_image("rgba8", 400, 400), [0.5, 0.5, 1.0, 1.0]))
    img, _ = run_once_with_output(c, pool, [], output)
    assert O.blockhash256(rgba(img)) in hashes()["solid"]
    assert np.array_equal(img.as_bytes(), O.solid(O.srgb_rgba8(400, 400), [0.5, 0.5, 1.0, 1.0]).data.reshape(-1))


@pytest.mark.parametrize("method", ["Scharr3", "Scharr3To4Bit", "Scharr3To8Bit", "Prewitt", "Sobel"])
def test_run_derivative(pool, images, method):  # tests/blend.rs:582-621
    bg, _ = images
    c = CommandBuffer()
    background = c.input(bg.descriptor())
    output, _ = c.output(c.derivative(background, Derivative(DerivativeMethod[method])))
    img, _ = run_once_with_output(c, pool, [(background, bg.key())], output)
    assert O.blockhash256(rgba(img)) in hashes()["derived_" + method]


def test_run_transmute_bytes_equal(pool, images, fixtures):  # tests/blend.rs:340-375
    bg, _ = images
    c = CommandBuffer()
    inp = c.input(bg.descriptor())
    output, fmt = c.output(c.transmute(inp, Descriptor.with_srgb_image("luma_a16", 512, 512)))
    img, _ = run_once_with_output(c, pool, [(inp, bg.key())], output)
    assert fmt.texel.bits == SampleBits.UInt16x2
    assert np.array_equal(img.as_bytes(), fixtures["background"].reshape(-1))
    a16 = (np.frombuffer(img.as_bytes().tobytes(), np.uint16).reshape(512, 512, 2).astype(np.int64) + 128) // 257
    view = np.stack([a16[..., 0]] * 3 + [a16[..., 1]], -1).astype(np.uint8)  # DynamicImage::get_pixel of a LumaA16 image
    assert O.blockhash256(view) in hashes()["transmute"]


def test_run_palette(pool, images, fixtures):  # tests/blend.rs:377-424 (goldens differ per device; checked against the oracle)
    bg, _ = images
    c = CommandBuffer()
    inp = c.input(bg.descriptor())
    ramp = c.bilinear(Descriptor.with_srgb_image("rgba8", 400, 400),
                      Bilinear([0, 0, 0, 1], [0.7, 0, 0, 1], [0, 0, 0, 1], [0, 0.7, 0, 1], [0, 0, 0, 1], [0.3, 0.3, 0, 1]))
    sampled = c.palette(inp, Palette(Z.ColorChannel.R, Z.ColorChannel.G), ramp)
    output, _ = c.output(sampled)
    img, _ = run_once_with_output(c, pool, [(inp, bg.key())], output)
    od = O.srgb_rgba8(400, 400)
    oramp = O.bilinear(od, ([0, 0, 0, 1], [0.7, 0, 0, 1], [0, 0, 0, 1], [0, 0.7, 0, 1], [0, 0, 0, 1], [0.3, 0.3, 0, 1]))
    exp = O.palette(oracle_image(bg.descriptor(), fixtures["background"]), oramp, [1, 0, 0, 0], [0, 1, 0, 0])
    got = rgba(img)
    assert np.mean(np.all(got == exp.data.reshape(400, 400, 4), axis=-1)) > 0.99  # pow in the ramp's sRGB pack: a few coordinates flip


def test_bilinear_knobs(pool):  # tests/knobs.rs: one Executable, five launches with a patched parameter block
    c = CommandBuffer()
    like = Descriptor.with_srgb_image("rgba8", 512, 512)
    result = c.with_knob().bilinear(like, Bilinear([0, 0, 1, 1], [1, 1, 1, 1], [0, 0, 1, 1], [1, 1, 1, 1]))
    output, _ = c.output(result)
    executable = Linker.from_included().compile(c).lower_to(Capabilities.from_device(next(pool.iter_devices())))
    knob = executable.query_knob(RegisterKnob(0, result))
    assert knob is not None
    for idx, (um, uM, vm, vM) in enumerate(R.KNOBS):
        data = Bilinear(um, uM, vm, vM).into_std430()
        img, _ = run_executable_with_output(executable, pool, [], output, [(knob, data)])
        assert O.blockhash256(rgba(img)) in hashes()["bilinear-knob-%d" % idx]


def test_errors_at_launch_and_step(pool, images):
    bg, fg = images
    c = CommandBuffer()
    background = c.input(bg.descriptor())
    output, _ = c.output(c.chromatic_adaptation(background, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D50))
    executable = Linker.from_included().compile(c).lower_to(Capabilities.from_device(next(pool.iter_devices())))
    env = executable.from_pool(pool)
    with pytest.raises(StartError):  # MismatchedDescriptor (run.rs:376-380)
        env.bind(background, fg.key())
    with pytest.raises(StartError):  # launching with an unbound input: StartError::MissingKey
        executable.launch(executable.from_pool(pool))
    env.bind(background, bg.key())
    ex = executable.launch(env)
    while ex.is_running():
        ex.step()
    with pytest.raises(StepError):  # StepError::ProgramEnd (run.rs:395-408)
        ex.step()
    ex.retire_gracefully(pool).finish()


# ---------------------------------------------------------------- planner: fusion keeps the bytes
def chain_c1(c, inp):
    lch = c.color_convert(inp, Color.Oklab, Texel(Z.Block.Pixel, SampleBits.UInt8x4, SampleParts.LchA))
    return c.color_convert(lch, Color.SRGB, Texel.new_u8(SampleParts.RgbA))


def test_fusion_c1_chain(pool, images, fixtures):
    """BASELINE config 1: sRGB8 -> Oklab (declared LchA u8 register) -> sRGB8.  Fused: one kernel; the
    LCh register is quantised in registers; bytes equal the unfused two-kernel run."""
    bg, _ = images
    res = {}
    for fuse in (_ffi.FUSE_EXACT, _ffi.FUSE_NONE, _ffi.FUSE_WIDE):
        c = CommandBuffer()
        inp = c.input(bg.descriptor())
        output, _ = c.output(chain_c1(c, inp))
        res[fuse] = run_once_with_output(c, pool, [(inp, bg.key())], output, fuse)
    assert res[_ffi.FUSE_EXACT][1] == 1 and res[_ffi.FUSE_NONE][1] == 2 and res[_ffi.FUSE_WIDE][1] == 1
    assert np.array_equal(res[_ffi.FUSE_EXACT][0].as_bytes(), res[_ffi.FUSE_NONE][0].as_bytes())
    # WIDE skips the 8-bit LCh quantisation: closer to the source image than the exact chain
    src = fixtures["background"].reshape(-1).astype(int)
    e_exact = np.abs(res[_ffi.FUSE_EXACT][0].as_bytes().astype(int) - src).mean()
    e_wide = np.abs(res[_ffi.FUSE_WIDE][0].as_bytes().astype(int) - src).mean()
    assert e_wide < 0.1 < e_exact


def test_fusion_into_composition(pool, images, fixtures):
    """convert(above) -> blend over convert(below) -> convert(result): the producer of `above` becomes
    source-side steps, the consumer becomes destination-side steps; same bytes as one kernel per op."""
    bg, fg = images
    res = {}
    for fuse in (_ffi.FUSE_EXACT, _ffi.FUSE_NONE):
        c = CommandBuffer()
        background, foreground = c.input(bg.descriptor()), c.input(fg.descriptor())
        a = c.color_convert(foreground, Color.Rgb(Z.Primaries.Bt2020, Transfer.Srgb), Texel.new_u8(SampleParts.RgbA))
        b = c.color_convert(background, Color.Rgb(Z.Primaries.Bt2020, Transfer.Srgb), Texel.new_u8(SampleParts.RgbA))
        blended = c.blend(b, Rectangle(100, 60, 257, 211), a, Blend.Alpha)
        back = c.color_convert(blended, Color.SRGB, Texel.new_u8(SampleParts.RgbA))
        output, _ = c.output(back)
        res[fuse] = run_once_with_output(c, pool, [(background, bg.key()), (foreground, fg.key())], output, fuse)
    assert res[_ffi.FUSE_NONE][1] == 4
    assert res[_ffi.FUSE_EXACT][1] == 2  # `below` must exist in memory: its producer stays a kernel
    assert np.array_equal(res[_ffi.FUSE_EXACT][0].as_bytes(), res[_ffi.FUSE_NONE][0].as_bytes())


def test_resize_modes_program(pool, images, fixtures):
    bg, _ = images
    for mode, omode in ((ResizeMode.Reference, "reference"), (ResizeMode.Nearest, "nearest"), (ResizeMode.Bilinear, "bilinear")):
        c = CommandBuffer()
        inp = c.input(bg.descriptor())
        output, fmt = c.output(c.resize(inp, (300, 200), mode))
        assert fmt.size() == (300, 200)
        img, n = run_once_with_output(c, pool, [(inp, bg.key())], output)
        assert n == 1
        exp = O.resize(oracle_image(bg.descriptor(), fixtures["background"]), (300, 200), omode)
        assert np.array_equal(img.as_bytes(), exp.data.reshape(-1))


def test_crop_quirk(pool, images, fixtures):  # command.rs:971-978: output keeps the source size
    bg, _ = images
    c = CommandBuffer()
    inp = c.input(bg.descriptor())
    output, fmt = c.output(c.crop(inp, Rectangle(100, 50, 356, 306)))
    assert fmt.size() == (512, 512)
    img, _ = run_once_with_output(c, pool, [(inp, bg.key())], output)
    exp = O.crop(oracle_image(bg.descriptor(), fixtures["background"]), (100, 50, 356, 306))
    assert np.array_equal(img.as_bytes(), exp.data.reshape(-1))


def test_run_swap(pool, images, fixtures):  # tests/blend.rs:426-447: extract R and G, inject them swapped
    bg, _ = images
    c = CommandBuffer()
    inp = c.input(bg.descriptor())
    channel_r = c.extract(inp, Z.ColorChannel.R)
    channel_g = c.extract(inp, Z.ColorChannel.G)
    assert c.describe_reg(channel_r).texel.parts == SampleParts.R and c.describe_reg(channel_r).texel.bits == SampleBits.UInt8
    intermediate = c.inject(inp, Z.ColorChannel.G, channel_r)
    swapped = c.inject(intermediate, Z.ColorChannel.R, channel_g)
    output, _ = c.output(swapped)
    img, _ = run_once_with_output(c, pool, [(inp, bg.key())], output)
    assert O.blockhash256(rgba(img)) in hashes()["swapped"]
    ob = oracle_image(bg.descriptor(), fixtures["background"])
    exp = O.inject(O.inject(ob, "G", O.extract(ob, "R")), "R", O.extract(ob, "G"))
    d = np.abs(img.as_bytes().astype(int) - exp.data.reshape(-1).astype(int))
    assert d.max() <= 1 and np.mean(d == 0) > 0.99  # the single-channel registers pack through pow (sRGB OETF)
    from zosimos_b200.command import CommandError
    with pytest.raises(CommandError):  # `above` must be a single-channel image of the matching texel
        c.inject(inp, Z.ColorChannel.R, inp)


def test_executable_reuse_cuda_graph(pool, images, fixtures):
    """SURVEY.md 8f-1 (tests/loop.rs, tests/knobs.rs): one Execution re-run many times.  The second and
    later runs go through ONE CUDA-graph launch; patching a knob re-captures; results always equal a
    fresh, stepped execution with the same knob value."""
    bg, fg = images
    c = CommandBuffer()
    background, foreground = c.input(bg.descriptor()), c.input(fg.descriptor())
    ramp = c.with_knob().bilinear(fg.descriptor(), Bilinear([0, 0, 1, 1], [1, 1, 1, 1], [0, 0, 1, 1], [1, 1, 1, 1]))
    rect = Rectangle.with_layout(fg.descriptor().layout)
    mixed = c.inscribe(background, rect, foreground)
    shifted = c.affine(mixed, Affine.new(AffineSample.Nearest).shift(100.0, 50.0), ramp)
    result = c.chromatic_adaptation(shifted, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D50)
    output, _ = c.output(result)
    caps = Capabilities.from_device(next(pool.iter_devices()), _ffi.FUSE_NONE)  # one kernel per op: a real multi-kernel schedule
    executable = Linker.from_included().compile(c).lower_to(caps)
    knob = executable.query_knob(RegisterKnob(0, ramp))
    binds = [(background, bg.key()), (foreground, fg.key())]
    datas = [Bilinear(um, uM, vm, vM).into_std430() for (um, uM, vm, vM) in R.KNOBS[:3]]
    fresh = [rgba(run_executable_with_output(executable, pool, binds, output, [(knob, d)])[0]) for d in datas]
    assert not np.array_equal(fresh[0], fresh[1])

    env = executable.from_pool(pool)
    env.knob(knob, datas[0])
    for reg, key in binds:
        env.bind(reg, key)
    ex = executable.launch(env)
    assert ex.kernel_count() > 2
    while ex.is_running():
        ex.step().block_on()

    def current():
        r = ex.retire_gracefully(pool)
        return rgba(r.output(output))

    assert np.array_equal(current(), fresh[0])
    ex.rerun().block_on()                      # run 1 of zos_program_run: eager
    assert np.array_equal(current(), fresh[0])
    for _ in range(3):                         # captured, then replayed
        ex.rerun().block_on()
    assert ex.graph_launches() == 3
    assert np.array_equal(current(), fresh[0])
    ex.rerun({knob: datas[1]}).block_on()      # knob patched: re-captured with the new parameter block
    assert np.array_equal(current(), fresh[1])
    ex.rerun({knob: datas[2]}, graph=False).block_on()
    assert np.array_equal(current(), fresh[2])
    ex.rerun().block_on()
    assert np.array_equal(current(), fresh[2])
    assert ex.graph_launches() == 5
    ex.retire_gracefully(pool).finish()


def _luma_as_rgba(img, bits16, alpha):
    d = img.descriptor()
    a = np.frombuffer(img.as_bytes().tobytes(), dtype=np.uint16 if bits16 else np.uint8).reshape(d.layout.height, d.layout.width, -1).astype(np.int64)
    if bits16:
        a = (a + 128) // 257
    lum = a[..., 0]
    al = a[..., 1] if alpha else np.full_like(lum, 255)
    return np.stack([lum, lum, lum, al], -1).astype(np.uint8)


@pytest.mark.parametrize("case", ["distribution_normal2d", "distribution_normal1d", "distribution_u8"])
def test_run_distribution_normal2d(pool, case):  # tests/blend.rs:227-312
    from zosimos_b200.command import DistributionNormal2d
    kind = "luma8" if case == "distribution_u8" else "luma_a16"
    desc = Descriptor.with_srgb_image(kind, 400, 400)
    dist = DistributionNormal2d.with_diagonal(0.2, 0.2) if case == "distribution_normal2d" else DistributionNormal2d.with_direction([0.04998, 0.0501])
    c = CommandBuffer()
    output, _ = c.output(c.distribution_normal2d(desc, dist))
    img, _ = run_once_with_output(c, pool, [], output)
    assert O.blockhash256(_luma_as_rgba(img, kind == "luma_a16", kind == "luma_a16")) in hashes()[case]
    # against the oracle: exp() differs in the last bits between the SFU path and libm, the 16-bit truncating pack may flip by a code
    params = O.normal2d_with_diagonal(0.2, 0.2) if case == "distribution_normal2d" else O.normal2d_with_direction(0.04998, 0.0501)
    assert np.allclose(dist.params, params, rtol=1e-6, atol=0)
    exp = O.distribution_normal2d(oracle_desc(desc), params)
    dt = np.uint16 if kind == "luma_a16" else np.uint8
    g = np.frombuffer(img.as_bytes().tobytes(), dtype=dt).astype(int); e = np.frombuffer(np.ascontiguousarray(exp.data).tobytes(), dtype=dt).astype(int)
    # the steep 1-d gaussian amplifies last-bit differences of the parameters (hypot / fma on the host) and of exp() by the
    # size of the exponent: a relative bound, plus the truncating pack's +-1
    assert (np.abs(g - e) <= 2 + 5e-4 * e).all()
    assert np.mean(np.abs(g - e) <= 1) > 0.9


def test_run_fractal_noise(pool):  # tests/blend.rs:314-338; integer hash + fixed operation order: bit exact
    from zosimos_b200.command import FractalNoise
    desc = Descriptor.with_srgb_image("rgba8", 400, 400)
    noise = FractalNoise.with_octaves(4)
    c = CommandBuffer()
    output, _ = c.output(c.distribution_fractal_noise(desc, noise))
    img, _ = run_once_with_output(c, pool, [], output)
    assert O.blockhash256(rgba(img)) in hashes()["distribution_fractal2d"]
    exp = O.distribution_fractal_noise(O.srgb_rgba8(400, 400), O.fractal_noise_with_octaves(4))
    assert np.array_equal(rgba(img), exp.data.reshape(400, 400, 4))
    noise.set_damping(0.5)
    c = CommandBuffer()
    output, _ = c.output(c.distribution_fractal_noise(desc, noise))
    img2, _ = run_once_with_output(c, pool, [], output)
    exp2 = O.distribution_fractal_noise(O.srgb_rgba8(400, 400), O.fractal_noise_with_octaves(4, 0.5))
    assert np.array_equal(rgba(img2), exp2.data.reshape(400, 400, 4))
    assert not np.array_equal(rgba(img2), rgba(img))


def _luma8_image(width=8, height=8):
    return Descriptor.with_srgb_image("luma8", width, height)


def test_run_from_buffer(pool):  # tests/buffer.rs:39-65: an 8x8 Luma8 image laid out with 256-byte rows in a byte buffer
    c = CommandBuffer()
    a = bytearray(b"\xff" * (8 * 256))
    a[256:264] = b"\x00" * 8
    buffer = c.buffer_init(bytes(a))
    result = c.from_buffer(buffer, _luma8_image())
    output, _ = c.output(result)
    img, _ = run_once_with_output(c, pool, [], output)
    got = img.as_bytes().reshape(8, 8)
    exp = np.full((8, 8), 255, np.uint8); exp[1, :] = 0
    assert np.array_equal(got, exp)
    lum = got.astype(np.uint8)
    assert O.blockhash256(np.stack([lum, lum, lum, np.full_like(lum, 255)], -1)) in hashes()["from_buffer"]


def test_run_from_buffer_knob(pool):  # tests/buffer.rs:67-119: the buffer's content is the knob
    c = CommandBuffer()
    a = bytearray(b"\xff" * (8 * 256))
    buffer = c.with_knob().buffer_init(bytes(a))
    result = c.from_buffer(buffer, _luma8_image())
    output, _ = c.output(result)
    executable = Linker.from_included().compile(c).lower_to(Capabilities.from_device(next(pool.iter_devices())))
    knob = executable.query_knob(RegisterKnob(0, buffer))
    assert knob is not None
    a[256:264] = b"\x00" * 8
    img, _ = run_executable_with_output(executable, pool, [], output, [(knob, bytes(a))])
    lum = img.as_bytes().reshape(8, 8)
    assert lum[1].max() == 0 and lum[0].min() == 255
    assert O.blockhash256(np.stack([lum, lum, lum, np.full_like(lum, 255)], -1)) in hashes()["from_buffer-with-knob"]
    img2, _ = run_executable_with_output(executable, pool, [], output)  # without the knob: the initial content
    assert img2.as_bytes().min() == 255


def test_run_bilinear_from_buffer(pool):  # tests/buffer.rs:121-149: the generator's parameter block is read from device memory
    c = CommandBuffer()
    desc = Descriptor.with_srgb_image("rgba8", 256, 256)
    params = np.asarray([[0, 0, 0, 1], [0, 0, 0.7, 1], [0, 0, 0.3, 1], [0, 1, 0.3, 1], [0, 0, 0, 1], [0, 0, 0, 1]], np.float32)
    buffer = c.buffer_init(params.tobytes())
    result = c.with_buffer(buffer).bilinear(desc, Bilinear([0] * 4, [0] * 4, [0] * 4, [0] * 4))
    output, _ = c.output(result)
    img, _ = run_once_with_output(c, pool, [], output)
    assert O.blockhash256(rgba(img)) in hashes()["bilinear_from_buffer"]
    exp = O.bilinear(O.srgb_rgba8(256, 256), params)
    assert np.array_equal(rgba(img), exp.data.reshape(256, 256, 4))
