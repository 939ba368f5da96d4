"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the headers declare,
and the C++ host mirror of the op builder (CommandBuffer / Linker) behaves like the reference:
same checks, same error kinds (lib/zosimos/src/command.rs), same parameter matrices as the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import zosimos_b200 as Z
from oracle import oracle as O
from zosimos_b200 import _ffi, command
from zosimos_b200.buffer import Color, Descriptor, SampleBits, SampleParts, Texel, Transfer
from zosimos_b200.command import (Affine, AffineSample, Bilinear, Blend, ChromaticAdaptationMethod, CommandBuffer, CommandError,
                                  CommandErrorKind, Derivative, DerivativeMethod, Linker, Rectangle, ResizeMode)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zosh?_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.lib()
    names = declared("zosimos_cuda.h") + declared("zosimos_host.h")
    assert len(names) > 60
    for n in names:
        assert hasattr(lib, n), n
    # and the ctypes table covers the device header completely
    assert set(declared("zosimos_cuda.h")) == set(_ffi.SIGNATURES)
    assert lib.zos_abi_version() == 5


def test_no_cpu_fallback():
    """Without a device the product refuses to run (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_ffi.ZosError) as e:
        Z.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_layout_helpers():
    lib = _ffi.lib()
    assert lib.zos_aligned_row_stride(157, 4) == 768 and lib.zos_aligned_row_stride(3840, 4) == 15360  # buffer.rs:121-134
    assert [lib.zos_bits_bytes(int(b)) for b in SampleBits] == [b.bytes() for b in SampleBits]
    d = Descriptor.with_srgb_image("rgba8", 10, 10).to_ffi()
    f = _ffi.ZosTexFmt()
    assert lib.zos_desc_texfmt(d, f) == 0 and f.storage == 1  # native Rgba8UnormSrgb (program.rs:794-805)
    d.transfer = int(Transfer.Bt709)
    assert lib.zos_desc_texfmt(d, f) == 0 and f.storage == 0  # staged
    d.bits = int(SampleBits.UInt8x3); d.texel_stride = 3
    assert lib.zos_desc_texfmt(d, f) == _ffi.ERR_UNSUPPORTED  # 3-byte texels: stage.rs:63-72
    d = Descriptor.with_srgb_image("rgba8", 10, 10).with_color(Color.Oklab).to_ffi()
    d.parts = int(SampleParts.LchA)
    assert lib.zos_desc_texfmt(d, f) == 0 and f.transfer == 0x100 and f.parts == int(SampleParts.LchA)  # program.rs:882-890


def test_color_matrices_match_oracle_bit_for_bit():
    for prim, name in ((Z.Primaries.Bt709, "bt709"), (Z.Primaries.Bt2020, "bt2020"), (Z.Primaries.Bt601_625, "bt601_625")):
        for wp in (Z.Whitepoint.D65, Z.Whitepoint.D50, Z.Whitepoint.E):
            got = command.to_xyz_matrix(prim, wp)
            assert np.array_equal(got, O.to_xyz(name, wp.name).astype(np.float32))
    for m, name in ((ChromaticAdaptationMethod.VonKries, "vonkries"), (ChromaticAdaptationMethod.BradfordVonKries, "bradford"),
                    (ChromaticAdaptationMethod.Xyz, "xyz")):
        got = command.adaptation_matrix(m, Z.Whitepoint.D65, Z.Whitepoint.D50)
        assert np.array_equal(got, O.adaptation_matrix(name, "D65", "D50").astype(np.float32))
    with pytest.raises(CommandError) as e:
        command.adaptation_matrix(ChromaticAdaptationMethod.BradfordNonLinear, Z.Whitepoint.D65, Z.Whitepoint.D50)
    assert e.value.kind == CommandErrorKind.Unimplemented  # command.rs:3327-3331


def test_rectangles():  # command.rs:3649-3658 + the normalize quirk (3536-3543)
    small, large = Rectangle.with_width_height(2, 2), Rectangle.with_width_height(4, 4)
    assert large == large.join(small) and small == large.meet(small)
    assert large.contains(small) and not small.contains(large)
    assert Rectangle(0, 0, 157, 151).normalize() == Rectangle(0, 0, 157, 157)


def test_affine_left_multiplication():
    from tests import refpipes as R
    a = Affine.new(AffineSample.Nearest).shift(-(157 // 2), -(151 // 2)).rotate(np.float32(np.pi) / np.float32(4)).shift(256, 256)
    assert np.allclose(np.array(a.transformation).reshape(3, 3), R.affine_matrix_blend_rs(157, 151, 512, 512), rtol=0, atol=1e-5)


def srgb(w, h):
    return Descriptor.with_srgb_image("rgba8", w, h)


def test_simple_program_compiles():  # command.rs:3660-3698
    cb = CommandBuffer()
    bg, fg = cb.input(srgb(512, 512)), cb.input(srgb(157, 151))
    r = cb.inscribe(bg, Rectangle(0, 0, 157, 151), fg)
    _, fmt = cb.output(r)
    assert fmt.layout == srgb(512, 512).layout
    prog = Linker.from_included().compile(cb)
    kinds = [o.kind for o in prog.ops()]
    assert kinds == [_ffi.OP_INPUT, _ffi.OP_INPUT, _ffi.OP_COMPOSE, _ffi.OP_OUTPUT]
    c = prog.ops()[2].compose
    assert list(c.tgt) == [0, 0, 157, 157] and list(c.sel) == [0, 0, 157, 151]  # the quirk reaches the kernel parameters


def test_builder_errors_match_reference():
    cb = CommandBuffer()
    bg, fg = cb.input(srgb(512, 512)), cb.input(srgb(157, 151))
    with pytest.raises(CommandError) as e:  # rect != layout of above: command.rs:1196-1198 -> OTHER
        cb.inscribe(bg, Rectangle(0, 0, 100, 100), fg)
    assert e.value.kind == CommandErrorKind.Other
    with pytest.raises(CommandError) as e:  # not contained: command.rs:1202-1206
        cb.inscribe(fg, Rectangle(0, 0, 512, 512), bg)
    assert e.value.kind == CommandErrorKind.Other
    other = cb.input(Descriptor.with_srgb_image("rgba16", 157, 151))
    with pytest.raises(CommandError) as e:  # chroma mismatch: command.rs:1186-1190
        cb.inscribe(bg, Rectangle(0, 0, 157, 151), other)
    assert e.value.kind == CommandErrorKind.ConflictingTypes and e.value.is_type_err()
    with pytest.raises(CommandError) as e:  # affine chroma mismatch -> TYPE_ERR (command.rs:1646-1648)
        cb.affine(bg, Affine.new(AffineSample.Nearest), other)
    assert e.value.kind == CommandErrorKind.GenericTypeError
    with pytest.raises(CommandError) as e:  # singular matrix (command.rs:1650-1657)
        cb.affine(bg, Affine(AffineSample.Nearest, [1, 0, 0, 2, 0, 0, 0, 0, 1]), fg)
    assert e.value.kind == CommandErrorKind.Other
    with pytest.raises(CommandError) as e:  # inconsistent input (command.rs:744-748)
        cb.input(Descriptor(Z.ByteLayout(4, 4, 16, 3), Color.SRGB, Texel.new_u8(SampleParts.RgbA)))
    assert e.value.kind == CommandErrorKind.BadDescriptor
    with pytest.raises(CommandError) as e:  # no conversion between different whitepoints (command.rs:1021, 1077-1084)
        cb.color_convert(bg, Color.Rgb(Z.Primaries.Bt709, Transfer.Srgb, Z.Whitepoint.D50), Texel.new_u8(SampleParts.RgbA))
    assert e.value.kind == CommandErrorKind.BadDescriptor
    with pytest.raises(CommandError) as e:  # transmute to another texel size (command.rs:1329-1336)
        cb.transmute(bg, Descriptor.with_srgb_image("rgba16", 512, 512))
    assert e.value.kind == CommandErrorKind.ConflictingTypes
    with pytest.raises(CommandError) as e:
        cb.transmute(bg, Descriptor.with_srgb_image("luma_a16", 256, 512))
    assert e.value.kind == CommandErrorKind.BadDescriptor
    with pytest.raises(CommandError) as e:  # Roberts & co: CompileError::NotYetImplemented (command.rs:3402-3408)
        cb.derivative(bg, Derivative(DerivativeMethod.Roberts))
    assert e.value.kind == CommandErrorKind.Unimplemented
    with pytest.raises(CommandError):
        cb.output(command.Register(99))
    # blend is implemented here (the reference returns UNIMPLEMENTED, command.rs:1510-1519)
    r = cb.blend(bg, Rectangle(10, 10, 167, 161), fg, Blend.Alpha)
    assert cb.describe_reg(r) == cb.describe_reg(bg)


def test_liveness_drops_dead_operations():  # command.rs:2216-2291
    cb = CommandBuffer()
    a = cb.input(srgb(64, 64))
    dead = cb.chromatic_adaptation(a, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D50)
    live = cb.resize(a, (32, 32), ResizeMode.Bilinear)
    cb.output(live)
    ops = Linker.from_included().compile(cb).ops()
    assert [o.reg for o in ops] == [a.index, live.index, live.index + 1] and dead.index not in [o.reg for o in ops]


def test_color_convert_parameters():
    cb = CommandBuffer()
    a = cb.input(srgb(8, 8))
    lch = cb.color_convert(a, Color.Oklab, Texel(Z.Block.Pixel, SampleBits.UInt8x4, SampleParts.LchA))
    back = cb.color_convert(lch, Color.SRGB, Texel.new_u8(SampleParts.RgbA))
    cb.output(back)
    ops = Linker.from_included().compile(cb).ops()
    enc, dec = ops[1], ops[2]
    assert enc.steps[0].kind == _ffi.STEP_OKLAB_ENC and dec.steps[0].kind == _ffi.STEP_OKLAB_DEC
    T = O.to_xyz("bt709", "D65")
    assert np.array_equal(np.array(list(enc.steps[0].m), np.float32), T.astype(np.float32).reshape(9))
    assert np.array_equal(np.array(list(dec.steps[0].m), np.float32), O.inv3(T).astype(np.float32).reshape(9))
    d = cb.describe_reg(lch)
    assert d.texel.parts == SampleParts.LchA and d.color.model == Z.ColorModel.Oklab and d.size() == (8, 8)


def test_srgb_encoder_bucket_tables_exact():
    """The two bucket tables of the exact sRGB8 encoder (texel.cuh) against the definition
    code(x) = #{k >= 1 : x >= thr[k]}: every float within 64 patterns of a rounding threshold, every
    1009th pattern of [0, 1], and the range ends.  (All 1 065 353 219 patterns were checked once with a
    C++ brute force, 0 mismatches; this is the regression guard that runs everywhere.)"""
    import ctypes as C
    from zosimos_b200 import _ffi
    lib = _ffi.lib()
    thr = np.zeros(260, np.float32); b1 = np.zeros(2048, np.uint32); b2 = np.zeros(2048, np.uint32)
    n1, n2 = C.c_uint32(), C.c_uint32()
    st = lib.zos_srgb_encoder_tables(thr.ctypes.data_as(C.POINTER(C.c_float)), b1.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n1),
                                     b2.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n2))
    assert st == 0 and n1.value == 1665 and n2.value == 646
    tb = thr[1:256].view(np.uint32).astype(np.int64)
    one = int(np.float32(1.0).view(np.uint32))
    pts = [np.arange(0, one + 3, 1009, dtype=np.int64), np.arange(0, 4096, dtype=np.int64), np.arange(one - 4096, one + 3, dtype=np.int64)]
    for t in tb:
        pts.append(np.arange(t - 64, t + 65, dtype=np.int64))
    u = np.unique(np.clip(np.concatenate(pts), 0, one + 2)).astype(np.uint32)
    x = u.view(np.float32)
    ref = np.searchsorted(thr[1:256], x, side="right").astype(np.uint32)
    # table 1: top 16 bits of max(bits, 2^-13) select the bucket
    idx = np.maximum(u, np.uint32(0x39000000))
    k1 = (idx >> 16) - 0x3900
    assert k1.max() < n1.value
    c1 = ((b1[k1] + idx) >> 16) & 0xff          # wraps mod 2^32 like the kernel
    assert np.array_equal(c1, ref)
    # table 2: the key comes from the bit pattern of x + 2^-5
    y = (x + np.float32(0.03125)).astype(np.float32)
    k2 = (y.view(np.uint32) >> 16) - 0x3d00
    assert k2.max() < n2.value
    c2 = (b2[k2] + idx) >> 24
    assert np.array_equal(c2, ref)


def test_buffer_builder_errors():
    """command.rs:937-968: from_buffer type and size checks; buffers are not image registers."""
    from zosimos_b200.buffer import Descriptor
    from zosimos_b200.command import CommandBuffer, CommandError
    c = CommandBuffer()
    img = c.input(Descriptor.with_srgb_image("rgba8", 16, 16))
    small = c.buffer_init(b"\x00" * 100)
    big = c.buffer_zero(8 * 256)
    assert c.buffer_size(big) == 8 * 256
    luma = Descriptor.with_srgb_image("luma8", 8, 8)
    with pytest.raises(CommandError):   # TYPE_ERR: an image register is not a buffer
        c.from_buffer(img, luma)
    with pytest.raises(CommandError):   # INVALID_CALL: 8 rows of 256 bytes do not fit 100 bytes
        c.from_buffer(small, luma)
    ok = c.from_buffer(big, luma)
    assert c.describe_reg(ok).layout.width == 8
    with pytest.raises(CommandError):   # a buffer register is not an image
        c.output(big)
    with pytest.raises(CommandError):
        c.with_buffer(img)
    with pytest.raises(CommandError):   # the bilinear block needs 96 bytes
        c.with_buffer(c.buffer_zero(32)).bilinear(Descriptor.with_srgb_image("rgba8", 4, 4))


# ---------------------------------------------------------------- functions and generics (tests/generic.rs; host.cpp)
def _op_bytes(op):
    return bytes(op)[:_ffi.ZosOp.data.offset]  # everything but the pointers to blobs / source text


def _same_ops(a, b):
    return len(a) == len(b) and all(_op_bytes(x) == _op_bytes(y) for x, y in zip(a, b))


def _palette_template(idx_desc, ramp):
    from zosimos_b200.command import GenericDeclaration, Palette
    t = CommandBuffer()
    var = t.generic(GenericDeclaration(bounds=()))
    img = t.input_generic(var)
    idx = t.bilinear(idx_desc, ramp)
    out, desc = t.output(t.palette(img, Palette(height=Z.ColorChannel.R, width=Z.ColorChannel.G), idx))
    assert desc is None  # the types of a template are bound by its caller
    return t


def test_generic_function_is_inlined_into_the_caller():  # tests/generic.rs
    from zosimos_b200.command import InvocationArguments, Palette
    ramp = Bilinear([0] * 4, [0] * 4, [0] * 4, [0] * 4, [0] * 4, [1, 1, 0, 0])
    idx_desc = Descriptor.with_texel(Texel.new_u8(SampleParts.RgbA), 64, 48)
    template = _palette_template(idx_desc, ramp)
    sig = template.computed_signature()
    assert (sig.num_generics, sig.num_inputs, sig.num_outputs) == (1, 1, 1)

    main = CommandBuffer()
    f = main.function(sig)
    inp = main.input(srgb(32, 32))
    ty = main.register_descriptor(inp)
    n0 = command.host_lib().zosh_cb_num_ops(main._h)
    with pytest.raises(CommandError) as e:  # wrong number of generics: INVALID_CALL
        main.invoke(f, InvocationArguments(generics=[], arguments=[inp]))
    assert e.value.is_type_err()
    with pytest.raises(CommandError) as e:  # the argument does not have the bound type
        main.invoke(f, InvocationArguments(generics=[idx_desc], arguments=[inp]))
    assert e.value.is_type_err()
    with pytest.raises(CommandError):       # unknown function variable: BAD_REGISTER
        main.invoke(command.FunctionVar(3), InvocationArguments(generics=[ty], arguments=[inp]))
    assert command.host_lib().zosh_cb_num_ops(main._h) == n0  # failed calls leave the caller untouched
    (res,) = main.invoke(f, InvocationArguments(generics=[ty], arguments=[inp]))
    _, out_desc = main.output(res)
    assert out_desc.size() == (64, 48) and out_desc.color == ty.color  # layout of the indices, chroma of the palette

    linker = Linker.from_included()
    with pytest.raises(CommandError):  # link tables must name the invoked function
        linker.link(main, [], [template], [[2], []])
    with pytest.raises(CommandError):  # one table per program
        linker.link(main, [], [template], [[1]])
    with pytest.raises(CommandError) as e:  # another template with the same shape is still another function
        linker.link(main, [], [_palette_template(idx_desc, ramp)], [[1], []])
    assert e.value.is_type_err()
    linked = linker.link(main, [], [template], [[1], []]).ops()

    direct = CommandBuffer()
    i2 = direct.input(srgb(32, 32))
    direct.output(direct.palette(i2, Palette(height=Z.ColorChannel.R, width=Z.ColorChannel.G), direct.bilinear(idx_desc, ramp)))
    assert _same_ops(linked, linker.compile(direct).ops())  # same stream as the pipeline written without the function


def test_function_with_two_results_called_twice():
    from zosimos_b200.command import GenericDeclaration, InvocationArguments
    t = CommandBuffer()
    var = t.generic(GenericDeclaration())
    fixed = t.input(srgb(16, 16))        # a concrete parameter next to the generic one
    img = t.input_generic(var)
    placed = t.inscribe(img, Rectangle(0, 0, 16, 16), fixed)
    t.output(placed)
    t.output(t.resize(placed, (8, 8), ResizeMode.Nearest))
    sig = t.computed_signature()
    assert (sig.num_generics, sig.num_inputs, sig.num_outputs) == (1, 2, 2)

    def build(with_function):
        c = CommandBuffer()
        small, a, b = c.input(srgb(16, 16)), c.input(srgb(64, 64)), c.input(srgb(40, 32))
        outs = []
        if with_function:
            f = c.function(sig)
            for big in (a, b):
                outs += c.invoke(f, InvocationArguments(generics=[c.register_descriptor(big)], arguments=[small, big]))
        else:
            for big in (a, b):
                p = c.inscribe(big, Rectangle(0, 0, 16, 16), small)
                outs += [p, c.resize(p, (8, 8), ResizeMode.Nearest)]
        sizes = [c.describe_reg(r).size() for r in outs]
        for r in outs:
            c.output(r)
        return c, sizes
    with_f, sizes = build(True)
    assert sizes == [(64, 64), (8, 8), (40, 32), (8, 8)]  # one monomorphic copy per call
    without, _ = build(False)
    linker = Linker.from_included()
    assert _same_ops(linker.link(with_f, [], [t], [[1], []]).ops(), linker.compile(without).ops())

    # a callee that does not type check under the bound types fails at invoke and is rolled back
    c = CommandBuffer()
    f = c.function(sig)
    small, tiny = c.input(srgb(16, 16)), c.input(srgb(8, 8))
    n0 = command.host_lib().zosh_cb_num_ops(c._h)
    with pytest.raises(CommandError):  # inscribe: the rectangle is not contained in an 8x8 image
        c.invoke(f, InvocationArguments(generics=[c.register_descriptor(tiny)], arguments=[small, tiny]))
    assert command.host_lib().zosh_cb_num_ops(c._h) == n0
    with pytest.raises(CommandError) as e:  # the concrete parameter is checked against its declared type
        c.invoke(f, InvocationArguments(generics=[c.register_descriptor(small)], arguments=[tiny, small]))
    assert e.value.is_type_err()


def test_template_restrictions():
    from zosimos_b200.command import GenericVar, InvocationArguments
    c = CommandBuffer()
    c.input(srgb(4, 4))
    with pytest.raises(CommandError):  # generics are declared before the first operation
        c.generic()
    with pytest.raises(CommandError) as e:  # command.rs:886-905: only generic command buffers have a computed signature here
        c.computed_signature()
    assert e.value.kind == CommandErrorKind.Unimplemented
    with pytest.raises(CommandError):
        c.input_generic(GenericVar(0))
    t = CommandBuffer()
    v = t.generic()
    with pytest.raises(CommandError):
        t.input_generic(GenericVar(v.index + 1))
    r = t.input_generic(v)
    with pytest.raises(CommandError):  # no descriptor before the types are bound
        t.describe_reg(r)
    with pytest.raises(CommandError):  # operands must be earlier entries of the record
        t.crop(command.Register(7), Rectangle(0, 0, 1, 1))
    t.output(r)
    with pytest.raises(CommandError) as e:  # a generic entry point cannot be compiled (CommandError::UNIMPLEMENTED)
        Linker.from_included().compile(t)
    assert e.value.kind == CommandErrorKind.Unimplemented
    with pytest.raises(CommandError):
        t.invoke(command.FunctionVar(0), InvocationArguments(generics=[], arguments=[]))


def test_plain_c_caller_of_the_host_api(tmp_path):
    """tests/c_abi/host_generic.c: include/zosimos_host.h compiles as C99 and a C program drives the op builder,
    invoke and link through the shared library (host code only, no GPU call)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    libdir = os.path.join(ROOT, "zosimos_b200")
    exe = str(tmp_path / "host_generic")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "host_generic.c"), "-o", exe, "-L", libdir, "-lzosimos_cuda",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "host_generic ok" in out.stdout, out.stderr


def _random_chain(cb, image, other, steps):
    """Apply `steps` (a list of small integers) as operations on `image`; `other` is a second 16 x 16 sRGB image."""
    from zosimos_b200.buffer import Whitepoint
    rgba8 = Texel.new_u8(SampleParts.RgbA)
    r = image
    for s in steps:
        if s == 0:
            r = cb.color_convert(r, Color.BT709_RGB, rgba8)
        elif s == 1:
            r = cb.color_convert(r, Color.Oklab, Texel(Z.Block.Pixel, SampleBits.UInt8x4, SampleParts.LchA))
            r = cb.color_convert(r, Color.SRGB, rgba8)
        elif s == 2:
            r = cb.resize(r, (24, 20), ResizeMode.Bilinear)
        elif s == 3:
            r = cb.resize(r, (33, 17), ResizeMode.Reference)
        elif s == 4:
            r = cb.inscribe(r, Rectangle(0, 0, 16, 16), other)
        elif s == 5:
            r = cb.blend(r, Rectangle(0, 0, 16, 16), other, Blend.Alpha)
        elif s == 6:
            r = cb.affine(r, Affine.new(AffineSample.Nearest).rotate(0.3).shift(2.0, 1.0), other)
        elif s == 7:
            r = cb.chromatic_adaptation(r, ChromaticAdaptationMethod.VonKries, Whitepoint.D50)
        elif s == 8:
            r = cb.derivative(r, Derivative(DerivativeMethod.Sobel, command.Direction.Width))
        elif s == 9:
            r = cb.inject(r, Z.ColorChannel.G, cb.extract(r, Z.ColorChannel.R))
        elif s == 10:
            r = cb.crop(r, Rectangle(1, 1, 9, 9))
        else:
            r = cb.with_knob().solid_rgba(cb.describe_reg(other) if not cb._is_template else srgb(16, 16), [0.25, 0.5, 0.75, 1.0])
    return r


def test_random_callees_inline_to_the_direct_stream():
    """Property: for random operation chains, a generic callee invoked from `main` links to exactly the stream of the same
    chain written into `main` directly (same operations, registers, descriptors and parameter blocks) -- or both fail."""
    from hypothesis import given, settings, strategies as st
    from zosimos_b200.command import InvocationArguments

    @settings(max_examples=60, deadline=None)
    @given(steps=st.lists(st.integers(0, 11), min_size=1, max_size=6), w=st.integers(16, 40), h=st.integers(16, 40))
    def prop(steps, w, h):
        t = CommandBuffer()
        var = t.generic()
        other_t, image_t = t.input(srgb(16, 16)), t.input_generic(var)
        t.output(_random_chain(t, image_t, other_t, steps))
        sig = t.computed_signature()

        direct = CommandBuffer()
        o, i = direct.input(srgb(16, 16)), direct.input(srgb(w, h))
        try:
            direct.output(_random_chain(direct, i, o, steps))
            expected = Linker.from_included().compile(direct).ops()
        except CommandError as e:
            expected = e.kind

        main = CommandBuffer()
        f = main.function(sig)
        o, i = main.input(srgb(16, 16)), main.input(srgb(w, h))
        try:
            (res,) = main.invoke(f, InvocationArguments(generics=[main.register_descriptor(i)], arguments=[o, i]))
            main.output(res)
            got = Linker.from_included().link(main, [], [t], [[1], []]).ops()
        except CommandError as e:
            got = e.kind
            assert command.host_lib().zosh_cb_num_ops(main._h) == 2  # rolled back to the two inputs
        if isinstance(expected, list):
            assert isinstance(got, list) and _same_ops(got, expected)
        else:
            assert got == expected
    prop()


def test_every_builder_replays():
    """The builders the random chains above do not reach: generators, byte buffers, transmute, palette, user operators."""
    from zosimos_b200.command import (DistributionNormal2d, FractalNoise, InvocationArguments, Palette, ShaderCommand)

    class Tint(ShaderCommand):
        def source(self):
            return "__device__ float4 zos_shade(float2 uv, const unsigned char* p, zos_tex a, zos_tex b) { return a.fetch(uv); }"

        def data(self, sd):
            sd.set_data(np.arange(4, dtype=np.float32))
            return srgb(20, 10)
    luma_a16 = Descriptor.with_srgb_image("luma_a16", 20, 10)
    ramp = Bilinear([0] * 4, [1, 0, 0, 1], [0] * 4, [0, 1, 0, 1], [0] * 4, [0] * 4)

    def body(cb, image):
        outs = [cb.transmute(image, luma_a16)]
        outs.append(cb.distribution_normal2d(srgb(20, 10), DistributionNormal2d.with_diagonal(0.1, 0.2)))
        outs.append(cb.distribution_fractal_noise(srgb(20, 10), FractalNoise.with_octaves(3)))
        buf = cb.buffer_init(np.asarray(ramp.flat(), np.float32).tobytes())
        assert cb.buffer_size(buf) == 96
        outs.append(cb.with_buffer(buf).bilinear(srgb(20, 10), ramp))
        zero = cb.buffer_zero(zos_bytes)
        outs.append(cb.from_buffer(zero, srgb(20, 10)))
        idx = cb.bilinear(srgb(20, 10), ramp)
        outs.append(cb.palette(image, Palette(width=Z.ColorChannel.R, height=Z.ColorChannel.G), idx))
        outs.append(cb.construct_dynamic(Tint()))
        outs.append(cb.unary_dynamic(image, Tint()))
        outs.append(cb.binary_dynamic(image, idx, Tint()))
        for r in outs:
            cb.output(r)
        return len(outs)
    zos_bytes = int(srgb(20, 10).to_aligned().row_stride) * 10

    t = CommandBuffer()
    n = body(t, t.input_generic(t.generic()))
    sig = t.computed_signature()
    assert sig.num_outputs == n
    main = CommandBuffer()
    f = main.function(sig)
    inp = main.input(srgb(20, 10))
    for r in main.invoke(f, InvocationArguments(generics=[main.register_descriptor(inp)], arguments=[inp])):
        main.output(r)
    linked_prog = Linker.from_included().link(main, [], [t], [[1], []])
    direct = CommandBuffer()
    body(direct, direct.input(srgb(20, 10)))
    direct_prog = Linker.from_included().compile(direct)
    a, b = linked_prog.ops(), direct_prog.ops()
    assert _same_ops(a, b)
    for x, y in zip(a, b):  # blobs and source text travel with the replayed operations
        assert x.data_len == y.data_len and (x.source or b"") == (y.source or b"")
        if x.data_len and x.data:
            assert C.string_at(x.data, x.data_len) == C.string_at(y.data, y.data_len)


def test_generic_function_calling_a_generic_function():
    """A template that declares and invokes another function, binding the callee's generic to its own: two levels of
    inlining give the stream of the flat pipeline; the inner type error surfaces at the outer invoke and rolls back."""
    from zosimos_b200.command import InvocationArguments
    rgba8 = Texel.new_u8(SampleParts.RgbA)

    inner = CommandBuffer()                      # inner<T>(small: srgb 16x16, image: T) -> inscribe
    v = inner.generic()
    small_i, image_i = inner.input(srgb(16, 16)), inner.input_generic(v)
    inner.output(inner.inscribe(image_i, Rectangle(0, 0, 16, 16), small_i))
    inner_sig = inner.computed_signature()

    outer = CommandBuffer()                      # outer<U>(image: U, small) -> (inner<U>(small, bt709(image)) resized, its input converted)
    u = outer.generic()
    f_inner = outer.function(inner_sig)
    image_o, small_o = outer.input_generic(u), outer.input(srgb(16, 16))
    conv = outer.color_convert(image_o, Color.SRGB, rgba8)
    (placed,) = outer.invoke(f_inner, InvocationArguments(generics=[u], arguments=[small_o, conv]))
    outer.output(outer.resize(placed, (8, 8), ResizeMode.Nearest))
    outer.output(conv)
    with pytest.raises(CommandError):            # counts are checked when the call is recorded
        outer.invoke(f_inner, InvocationArguments(generics=[], arguments=[small_o, conv]))
    with pytest.raises(CommandError):            # a generic the template does not have
        outer.invoke(f_inner, InvocationArguments(generics=[command.GenericVar(5)], arguments=[small_o, conv]))
    outer_sig = outer.computed_signature()
    assert (outer_sig.num_generics, outer_sig.num_inputs, outer_sig.num_outputs) == (1, 2, 2)

    main = CommandBuffer()
    f = main.function(outer_sig)
    big, small, tiny = main.input(srgb(48, 40)), main.input(srgb(16, 16)), main.input(srgb(8, 8))
    n0 = command.host_lib().zosh_cb_num_ops(main._h)
    with pytest.raises(CommandError):            # 16x16 does not fit into 8x8: the INNER inscribe fails, everything is undone
        main.invoke(f, InvocationArguments(generics=[main.register_descriptor(tiny)], arguments=[tiny, small]))
    assert command.host_lib().zosh_cb_num_ops(main._h) == n0
    a, b = main.invoke(f, InvocationArguments(generics=[main.register_descriptor(big)], arguments=[big, small]))
    assert main.describe_reg(a).size() == (8, 8) and main.describe_reg(b).size() == (48, 40)
    main.output(a); main.output(b)
    linker = Linker.from_included()
    with pytest.raises(CommandError):            # outer's own link table is checked as well: it calls program 2, not itself
        linker.link(main, [], [outer, inner], [[1], [1], []])
    linked = linker.link(main, [], [outer, inner], [[1], [2], []])

    flat = CommandBuffer()
    big, small, tiny = flat.input(srgb(48, 40)), flat.input(srgb(16, 16)), flat.input(srgb(8, 8))
    conv = flat.color_convert(big, Color.SRGB, rgba8)
    flat.output(flat.resize(flat.inscribe(conv, Rectangle(0, 0, 16, 16), small), (8, 8), ResizeMode.Nearest))
    flat.output(conv)
    assert _same_ops(linked.ops(), linker.compile(flat).ops())


def test_builder_fuzz_never_crashes():
    """Random (mostly invalid) arguments through the raw C entry points: every call answers ZOSH_OK or one of the error
    kinds, registers stay dense, and whatever was built still compiles."""
    from hypothesis import given, settings, strategies as st
    L = command.host_lib()
    reg = st.integers(-3, 12)
    u32 = st.one_of(st.integers(0, 24), st.sampled_from([0xFFFFFFFF, 0x80000000, 255, 256, 65535]))
    dim = st.one_of(st.integers(0, 70), st.sampled_from([0xFFFFFFFF, 1 << 16, 1 << 30]))

    def desc(draw):
        d = _ffi.ZosDesc()
        d.width, d.height = draw(dim), draw(dim)
        d.block, d.bits, d.parts, d.color, d.transfer = draw(st.integers(0, 3)), draw(u32), draw(u32), draw(st.integers(0, 5)), draw(u32)
        d.primaries, d.whitepoint, d.texel_stride = draw(st.integers(0, 7)), draw(st.integers(0, 12)), draw(st.integers(0, 17))
        d.row_stride = draw(st.integers(0, 1 << 20))
        return d

    @settings(max_examples=150, deadline=None)
    @given(data=st.data())
    def prop(data):
        draw = data.draw
        cb = L.zosh_cb_new()
        template = draw(st.booleans()) and draw(st.booleans())
        out = C.c_int32(-1)
        try:
            if template:
                assert L.zosh_cb_generic(cb, C.byref(out)) == 0
                L.zosh_cb_input_generic(cb, draw(st.integers(-1, 2)), C.byref(out))
            good = srgb(32, 24).to_ffi()
            assert L.zosh_cb_input(cb, C.byref(good), C.byref(out)) == 0
            for _ in range(draw(st.integers(1, 12))):
                k = draw(st.integers(0, 15))
                d = desc(draw) if draw(st.booleans()) else srgb(draw(st.integers(1, 40)), draw(st.integers(1, 40))).to_ffi()
                rect = command.ZoshRect(draw(dim), draw(dim), draw(dim), draw(dim))
                f24 = (C.c_float * 24)(*[draw(st.floats(-4, 4, width=32)) for _ in range(24)])
                a, b = draw(reg), draw(reg)
                st_ = [
                    lambda: L.zosh_cb_input(cb, C.byref(d), C.byref(out)),
                    lambda: L.zosh_cb_output(cb, a, C.byref(out)),
                    lambda: L.zosh_cb_color_convert(cb, a, C.byref(d), C.byref(out)),
                    lambda: L.zosh_cb_chromatic_adaptation(cb, a, draw(u32), draw(u32), C.byref(out)),
                    lambda: L.zosh_cb_inscribe(cb, a, rect, b, C.byref(out)),
                    lambda: L.zosh_cb_blend(cb, a, rect, b, draw(st.integers(-3, 14)), C.byref(out)),
                    lambda: L.zosh_cb_crop(cb, a, rect, C.byref(out)),
                    lambda: L.zosh_cb_affine(cb, a, f24, draw(u32), b, C.byref(out)),
                    lambda: L.zosh_cb_resize(cb, a, draw(dim), draw(dim), draw(u32), C.byref(out)),
                    lambda: L.zosh_cb_transmute(cb, a, C.byref(d), C.byref(out)),
                    lambda: L.zosh_cb_bilinear(cb, C.byref(d), f24, C.byref(out)),
                    lambda: L.zosh_cb_derivative(cb, a, draw(u32), draw(u32), C.byref(out)),
                    lambda: L.zosh_cb_palette(cb, a, b, f24, f24, C.byref(out)),
                    lambda: L.zosh_cb_extract(cb, a, draw(u32), C.byref(out)),
                    lambda: L.zosh_cb_inject(cb, a, draw(u32), b, C.byref(out)),
                    lambda: L.zosh_cb_from_buffer(cb, a, C.byref(d), C.byref(out)),
                ][k]()
                assert 0 <= st_ <= 6
                got = _ffi.ZosDesc()
                L.zosh_cb_describe(cb, draw(reg), C.byref(got))
            prog = command._P()
            st_ = L.zosh_compile(cb, C.byref(prog))
            assert st_ == (5 if template else 0)
            if st_ == 0:
                n = L.zosh_program_num_ops(prog)
                ops = L.zosh_program_ops(prog)
                for i in range(n):
                    assert all(s < ops[i].reg or s == -1 for s in ops[i].src)  # operands precede their users
                L.zosh_program_free(prog)
        finally:
            L.zosh_cb_free(cb)
    prop()


def test_descriptor_helpers_fuzz():
    """zos_desc_texfmt / zos_desc_device_bytes / zos_aligned_row_stride are host code: random descriptors either get a
    texture format or ZOS_ERR_UNSUPPORTED / INVALID, never a crash; strides are 256-aligned and cover the row."""
    from hypothesis import given, settings, strategies as st
    L = _ffi.lib()

    @settings(max_examples=300, deadline=None)
    @given(w=st.integers(0, 1 << 20), h=st.integers(0, 1 << 16), block=st.integers(0, 4), bits=st.integers(0, 40), parts=st.integers(0, 40),
           color=st.integers(0, 6), transfer=st.one_of(st.integers(0, 14), st.just(0x100)), ts=st.integers(0, 20))
    def prop(w, h, block, bits, parts, color, transfer, ts):
        d = _ffi.ZosDesc()
        d.width, d.height, d.block, d.bits, d.parts, d.color, d.transfer, d.texel_stride = w, h, block, bits, parts, color, transfer, ts
        f = _ffi.ZosTexFmt()
        st_ = L.zos_desc_texfmt(C.byref(d), C.byref(f))
        assert 0 <= st_ < len(_ffi.STATUS_NAMES)
        L.zos_desc_device_bytes(C.byref(d))
        nb = L.zos_bits_bytes(bits)
        assert 0 <= nb <= 16
        stride = L.zos_aligned_row_stride(w, ts)
        assert stride % 256 == 0 and stride >= w * ts and stride < w * ts + 256
    prop()


def test_every_entry_point_survives_null_arguments():
    """INTEGRATION.md's contract: nothing aborts across the boundary.  Every exported function of both headers is called
    with NULL / zero arguments in a child process (a crash there is a failure here); status-returning ones answer an error."""
    import subprocess
    import sys
    code = '''
import ctypes as C, sys
sys.path.insert(0, %r)
from zosimos_b200 import _ffi, command
L = _ffi.lib(); command.host_lib()
sigs = dict(_ffi.SIGNATURES); sigs.update(command._HOST_SIGNATURES)
def zero(t):
    if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents"): return None
    if issubclass(t, C.Structure): return t()
    if t in (C.c_float, C.c_double): return 0.0
    return 0
for name in sorted(sigs):
    res, args = sigs[name]
    print(name, flush=True)
    r = getattr(L, name)(*[zero(a) for a in args])
    if res is C.c_int32 and any(a is C.c_void_p or hasattr(a, "contents") for a in args) and name not in ("zos_ctx_device", "zosh_cb_with_knob", "zos_srgb_encoder_tables"):
        assert r != 0, name
print("ALL-OK")
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("ALL-OK"), (out.stdout[-300:], out.stderr[-1500:])


def test_colour_science_known_answers():
    """The third-party arithmetic of the path (image-canvas `to_xyz_row_matrix`, palette `TransformMatrix`) is
    not vendored in the reference, so it is pinned to the published tables both crates follow: Lindbloom's sRGB -> XYZ (D65)
    matrix and his Bradford / von Kries D65 -> D50 adaptation matrices (ASTM E308 whitepoints), and BT.2020's NPM
    (ITU-R BT.2020-2 / SMPTE RP 177, whose D65 is the chromaticity (0.3127, 0.3290): 3e-4 from the ASTM XYZ triple)."""
    from zosimos_b200.buffer import Primaries, Whitepoint
    srgb_d65 = [[0.4124564, 0.3575761, 0.1804375], [0.2126729, 0.7151522, 0.0721750], [0.0193339, 0.1191920, 0.9503041]]
    assert np.abs(command.to_xyz_matrix(Primaries.Bt709, Whitepoint.D65) - srgb_d65).max() < 2e-7
    bt2020 = [[0.6369580, 0.1446169, 0.1688810], [0.2627002, 0.6779981, 0.0593017], [0.0, 0.0280727, 1.0609851]]
    assert np.abs(command.to_xyz_matrix(Primaries.Bt2020, Whitepoint.D65) - bt2020).max() < 3e-4
    bradford = [[1.0478112, 0.0228866, -0.0501270], [0.0295424, 0.9904844, -0.0170491], [-0.0092345, 0.0150436, 0.7521316]]
    got = command.adaptation_matrix(ChromaticAdaptationMethod.BradfordVonKries, Whitepoint.D65, Whitepoint.D50)
    assert np.abs(got - bradford).max() < 2e-7
    von_kries = [[1.0160803, 0.0552297, -0.0521326], [0.0060666, 0.9955661, -0.0012235], [0.0, 0.0, 0.7578869]]
    got = command.adaptation_matrix(ChromaticAdaptationMethod.VonKries, Whitepoint.D65, Whitepoint.D50)
    assert np.abs(got - von_kries).max() < 2e-7
    # white is a fixed point: primaries -> XYZ of (1, 1, 1) is the whitepoint, and adaptation maps whitepoint to whitepoint
    for wp, xyz in ((Whitepoint.D65, (0.95047, 1.0, 1.08883)), (Whitepoint.D50, (0.96422, 1.0, 0.82521))):
        for prim in (Primaries.Bt709, Primaries.Bt2020, Primaries.Bt601_625):
            assert np.abs(command.to_xyz_matrix(prim, wp) @ np.ones(3, np.float32) - xyz).max() < 1e-6
    for method in (ChromaticAdaptationMethod.BradfordVonKries, ChromaticAdaptationMethod.VonKries):
        m = command.adaptation_matrix(method, Whitepoint.D65, Whitepoint.D50)
        assert np.abs(m @ np.array([0.95047, 1.0, 1.08883], np.float32) - (0.96422, 1.0, 0.82521)).max() < 1e-6


def test_generic_entry_point():
    """Linker::link with `tys` (command.rs:2083-2185): a generic `main` is compiled as its monomorphic copy under the bound
    types; its registers are translated for binding inputs and retiring outputs."""
    from zosimos_b200.command import InvocationArguments
    from zosimos_b200.program import Environment, Executable, Pool, StartError
    rgba8 = Texel.new_u8(SampleParts.RgbA)

    helper = CommandBuffer()                     # helper<T>(x: T) -> BT.709 transfer
    hv = helper.generic()
    helper.output(helper.color_convert(helper.input_generic(hv), Color.BT709_RGB, rgba8))
    helper_sig = helper.computed_signature()

    main = CommandBuffer()                       # main<T>(image: T, small: srgb 16x16): calls helper<T>, then inscribes
    t = main.generic()
    f = main.function(helper_sig)
    image, small = main.input_generic(t), main.input(srgb(16, 16))
    (conv,) = main.invoke(f, InvocationArguments(generics=[t], arguments=[image]))
    placed = main.inscribe(conv, Rectangle(0, 0, 16, 16), main.color_convert(small, Color.BT709_RGB, rgba8))
    out, _ = main.output(placed)

    linker = Linker.from_included()
    with pytest.raises(CommandError) as e:       # one type per generic
        linker.link(main, [], [helper], [[1], []])
    assert e.value.is_type_err()
    with pytest.raises(CommandError):            # 16x16 does not fit: the entry point does not type check under this type
        linker.link(main, [srgb(8, 8)], [helper], [[1], []])
    prog = linker.link(main, [srgb(40, 30)], [helper], [[1], []])

    flat = CommandBuffer()
    i2, s2 = flat.input(srgb(40, 30)), flat.input(srgb(16, 16))
    c2 = flat.color_convert(i2, Color.BT709_RGB, rgba8)
    o2, _ = flat.output(flat.inscribe(c2, Rectangle(0, 0, 16, 16), flat.color_convert(s2, Color.BT709_RGB, rgba8)))
    plain = linker.compile(flat)
    assert _same_ops(prog.ops(), plain.ops())
    assert [prog.register_index(r.index) for r in (image, small, out)] == [i2.index, s2.index, o2.index]
    assert prog.register_index(99) == -1 and [plain.register_index(r.index) for r in (i2, s2, o2)] == [i2.index, s2.index, o2.index]

    # binding goes through the translation (host-side part of Environment, no device involved)
    pool = Pool()
    big = pool.insert(srgb(40, 30), np.zeros(40 * 30 * 4, np.uint8))
    little = pool.insert(srgb(16, 16), np.zeros(16 * 16 * 4, np.uint8))
    env = Environment(Executable(prog, None), pool, None)
    env.bind(image, big.key()); env.bind(small, little.key())
    assert sorted(env.inputs) == [i2.index, s2.index]
    with pytest.raises(StartError):
        env.bind(small, big.key())               # MismatchedDescriptor
    with pytest.raises(StartError):
        env.bind(out, big.key())                 # not an input


def test_knobs_inside_invoked_functions_keep_their_register():
    """RegisterKnob{link_idx, register} -> Knob (command.rs:701-705, 2134-2145).  Knob ids count in emission order, the
    knobs inside invoked functions included, so a generic entry point's own knobs do NOT carry the numbers its template
    handed out: `query_knob` has to answer with the id of the operation the register became."""
    from zosimos_b200.command import InvocationArguments, RegisterKnob
    from zosimos_b200.program import Executable
    rgba8 = Texel.new_u8(SampleParts.RgbA)

    helper = CommandBuffer()                     # helper<T>(x: T): a knob-able conversion
    hv = helper.generic()
    h_conv = helper.with_knob().color_convert(helper.input_generic(hv), Color.BT709_RGB, rgba8)
    helper.output(h_conv)
    helper_sig = helper.computed_signature()

    main = CommandBuffer()                       # main<T>(image: T): helper<T>(image), then a knob-able adaptation
    t = main.generic()
    f = main.function(helper_sig)
    image = main.input_generic(t)
    (conv,) = main.invoke(f, InvocationArguments(generics=[t], arguments=[image]))
    adapted = main.with_knob().chromatic_adaptation(conv, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D50)
    out, _ = main.output(adapted)

    prog = Linker.from_included().link(main, [srgb(40, 30)], [helper], [[1], []])
    ops = prog.ops()
    by_knob = {o.knob: o for o in ops if o.knob}
    assert sorted(by_knob) == [1, 2]
    exe = Executable(prog, None)
    k_main = exe.query_knob(RegisterKnob(0, adapted))
    k_helper = exe.query_knob(RegisterKnob(1, h_conv))
    assert k_main is not None and k_helper is not None and k_main.index != k_helper.index
    # the knob of main's op is the one on the op main's register translates to; same for the helper's
    assert by_knob[k_main.index].reg == prog.register_index(adapted.index)
    assert by_knob[k_helper.index].dst == ops[[o.reg for o in ops].index(prog.register_index(adapted.index))].src[0]
    assert exe.query_knob(RegisterKnob(0, image)) is None and exe.query_knob(RegisterKnob(2, h_conv)) is None

    # a non-generic main: its own registers under link 0, the callee's under its link index
    flat = CommandBuffer()
    g = flat.function(helper_sig)
    i2 = flat.input(srgb(40, 30))
    (c2,) = flat.invoke(g, InvocationArguments(generics=[srgb(40, 30)], arguments=[i2]))
    a2 = flat.with_knob().chromatic_adaptation(c2, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D50)
    flat.output(a2)
    p2 = Linker.from_included().link(flat, [], [helper], [[1], []])
    e2 = Executable(p2, None)
    assert e2.query_knob(RegisterKnob(0, a2)).index == 2 and e2.query_knob(RegisterKnob(1, h_conv)).index == 1
    assert {o.knob: o.reg for o in p2.ops() if o.knob} == {1: c2.index, 2: a2.index}


def test_failed_builder_does_not_leave_a_pending_knob():
    """with_knob() marks the NEXT operation (command.rs:1865-1874); when that operation is rejected the mark is gone on both
    sides of the veneer instead of landing on an unrelated later call."""
    rgba8 = Texel.new_u8(SampleParts.RgbA)
    cb = CommandBuffer()
    i = cb.input(srgb(20, 20))
    with pytest.raises(CommandError):
        cb.with_knob().chromatic_adaptation(i, ChromaticAdaptationMethod.BradfordNonLinear, Z.Whitepoint.D50)  # Unimplemented
    conv = cb.color_convert(i, Color.BT709_RGB, rgba8)
    cb.output(conv)
    assert cb._knobs == {} and cb._pending_knob == 0
    assert all(o.knob == 0 for o in Linker.from_included().compile(cb).ops())
    with pytest.raises(CommandError):
        cb.with_knob().inscribe(i, Rectangle(0, 0, 99, 99), conv)
    later = cb.with_knob().solid_rgba(srgb(4, 4), [0, 0, 0, 1])
    assert cb._knobs == {later.index: 3}  # ids keep counting: the two failed calls consumed 1 and 2 without attaching them anywhere


def test_normal2d_constructors_match_the_reference_values():
    """shaders/distribution_normal2d.rs:25-100: the host layer's parameter blocks equal the oracle's, and with_direction
    carries 2 pi length^2 (0.031466 for the direction the reference tests use), not length^2 (0.005008)."""
    from zosimos_b200.command import DistributionNormal2d
    d = DistributionNormal2d.with_direction([0.04998, 0.0501])
    assert np.allclose(d.params, O.normal2d_with_direction(0.04998, 0.0501), rtol=1e-6, atol=0)
    assert abs(d.params[6] - 0.031466) < 1e-6
    assert np.frombuffer(d.into_std430(), np.float32)[6] == np.float32(d.params[6])
    g = DistributionNormal2d.with_diagonal(0.2, 0.2)
    assert np.allclose(g.params, O.normal2d_with_diagonal(0.2, 0.2), rtol=1e-6, atol=0)


def test_affine_box_width_minimises_modelled_bank_conflicts():
    """k_affine_f16 sizes its staged source box so that the taps of a 32-lane row meet as few shared-memory bank conflicts as possible
    (affine_f16.cu pick_box_width, host code).  An independent model of the banks -- LDS.64, 16 lanes per wavefront, bank pair =
    (column + row * width) mod 16, distinct addresses on one bank pair serialise -- must agree that the chosen width is the cheapest of
    the candidates, and the choice must repair the rotation the fixed round-1 rule (width = 2 mod 4) handled worst."""
    import math

    def cost(width, dx, dy):
        total = 0
        for ph in range(16):
            x0, y0 = 64.0 + 0.25 * (ph & 3) + 0.125, 64.0 + 0.25 * (ph >> 2) + 0.125
            for tap in range(4):
                for half in range(2):
                    addr = []
                    for l in range(16):
                        i = half * 16 + l
                        X = math.floor(np.float32(x0) + np.float32(i) * np.float32(dx) - np.float32(0.5)) + (tap & 1)
                        Y = math.floor(np.float32(y0) + np.float32(i) * np.float32(dy) - np.float32(0.5)) + (tap >> 1)
                        addr.append(Y * width + X)
                    total += max(len({a for a in addr if a % 16 == b}) for b in range(16))
        return total

    lib = _ffi.lib()
    assert lib.zos_affine_box_width(0, 1.0, 0.0) == -1 and lib.zos_affine_box_width(300, 1.0, 0.0) == -1
    for deg in (30.0, -30.0, 17.0, 45.0, 5.0, 90.0, 133.0):
        c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
        min_w = int(math.ceil(31 * (abs(c) + abs(s)))) + 5
        min_w += min_w & 1
        w = lib.zos_affine_box_width(min_w, c, s)
        cands = list(range(min_w, min_w + 16, 2))
        costs = {k: cost(k, c, s) for k in cands}
        assert w in cands and costs[w] == min(costs.values()), (deg, w, costs)
        assert w == min(k for k in cands if costs[k] == costs[w])  # ties: the narrowest
    # -30 degrees: the round-1 width (50) put 16 lanes on 2-3 bank pairs
    w = lib.zos_affine_box_width(48, math.cos(math.radians(-30.0)), math.sin(math.radians(-30.0)))
    assert cost(50, 0.8660254, -0.5) > 2.5 * cost(w, 0.8660254, -0.5)
