import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import raytracinginrust_b200 as rt
for name in ("cornell", "final"):
    hs = rt.HostScene(name)
    dev = rt.DeviceScene(hs.scene_desc)
    a, sa = dev.render(hs.camera, 40, 28, 6, 30, rt.render_opts(seed=3, integrator=hs.integrator, flags=rt._abi.FLAG_MEGAKERNEL))
    b, sb = dev.render(hs.camera, 40, 28, 6, 30, rt.render_opts(seed=3, integrator=hs.integrator, flags=rt._abi.FLAG_WAVEFRONT))
    print(name, "identical", np.array_equal(a, b, equal_nan=True), sa.paths, sb.paths, sa.rays, sb.rays, sb.kernel_launches)
