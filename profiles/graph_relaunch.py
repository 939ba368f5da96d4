"""profiles/graph_relaunch.py -- what relaunching a planned program as ONE CUDA graph buys.
A launch-bound program (small images, one kernel per op) is run N times eagerly and N times through
zos_program_run(ZOS_RUN_GRAPH); wall time per run, host side included, synchronised at the end only."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import zosimos_b200 as Z  # noqa: E402
from zosimos_b200 import _ffi  # noqa: E402
from zosimos_b200.buffer import Descriptor  # noqa: E402
from zosimos_b200.command import Affine, AffineSample, Bilinear, ChromaticAdaptationMethod, CommandBuffer, Linker, Rectangle  # noqa: E402
from zosimos_b200.program import Capabilities, Pool  # noqa: E402


def main():
    pool = Pool(); pool.request_device(0)
    rng = np.random.default_rng(0)
    for size in (128, 512, 2048):
        d = Descriptor.with_srgb_image("rgba8", size, size)
        a = pool.insert(d, rng.integers(0, 256, size * size * 4, dtype=np.uint8))
        b = pool.insert(d, rng.integers(0, 256, size * size * 4, dtype=np.uint8))
        c = CommandBuffer()
        ia, ib = c.input(d), c.input(d)
        x = c.inscribe(ia, Rectangle.with_layout(d.layout), ib)
        x = c.chromatic_adaptation(x, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D50)
        x = c.chromatic_adaptation(x, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D65)
        x = c.affine(x, Affine.new(AffineSample.Nearest).shift(3.0, 5.0), c.bilinear(d, Bilinear([0, 0, 1, 1], [1, 1, 1, 1], [0, 0, 1, 1], [1, 1, 1, 1])))
        x = c.chromatic_adaptation(x, ChromaticAdaptationMethod.Xyz, Z.Whitepoint.D50)
        out, _ = c.output(x)
        exe = Linker.from_included().compile(c).lower_to(Capabilities.from_device(next(pool.iter_devices()), _ffi.FUSE_NONE))
        env = exe.from_pool(pool); env.bind(ia, a.key()); env.bind(ib, b.key())
        ex = exe.launch(env)
        while ex.is_running():
            ex.step().block_on()
        n = 300
        res = {}
        for graph in (False, True):
            for _ in range(5):
                ex.rerun(graph=graph)
            ex.ctx.sync()
            t0 = time.perf_counter()
            for _ in range(n):
                ex.rerun(graph=graph)
            ex.ctx.sync()
            res[graph] = (time.perf_counter() - t0) / n * 1e6
        print("%4dx%-4d  %d kernels/run   eager %7.1f us/run   graph %7.1f us/run   x%.2f" % (size, size, ex.kernel_count(), res[False], res[True], res[False] / res[True]))
        ex.retire_gracefully(pool).finish()


if __name__ == "__main__":
    main()
