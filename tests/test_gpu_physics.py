"""GPU tier: the CUDA path against closed-form radiometry (the scenes of test_oracle_physics.py), through the C ABI.
The per-path parity tests compare the device with the oracle; these compare it with physics directly."""
import math

import numpy as np
import pytest

from test_oracle_physics import centre_paths, rect_form_factor

pytestmark = pytest.mark.gpu


def test_rect_and_sphere_lights_and_fog_on_device(rt, orc):
    A = rt._abi
    n = 60000
    px, py, smp = centre_paths(n, 3, 3)
    cam = rt.camera_new((6.0, 1.0, 0.0), (0.0, 0.0, 0.0), (0, 1, 0), 0.05, 1.0, 0.0, 6.0)

    def floor_under(light_of):
        b = rt.SceneBuilder()
        rho = 0.6
        floor = b.rect(A.PLANE_XZ, -500, 500, -500, 500, 0.0, b.lambertian(b.constant_texture((rho, rho, rho))))
        light = light_of(b)
        return rho, b.finish(b.list([floor, light]), b.list([light]))

    # rect light: rho * E * form factor (rect.rs:91-111 through the mixture pdf, main.rs:92-98)
    E, a, c, h = 5.0, 3.0, 2.0, 2.5
    rho, sd = floor_under(lambda b: b.flip(b.rect(A.PLANE_XZ, -a / 2, a / 2, -c / 2, c / 2, h, b.diffuse_light(b.constant_texture((E, E, E))))))
    dev = rt.DeviceScene(sd, device=0)
    rgb, _ = dev.path_radiance(cam, 3, 3, 50, rt.render_opts(seed=3, integrator=rt.INTEGRATOR_HEAD), px, py, smp)
    expect = rho * E * rect_form_factor(a, c, h)
    mean, sem = rgb[:, 0].mean(), rgb[:, 0].std() / math.sqrt(n)
    print("rect light: L = %.5f +- %.5f, closed form %.5f" % (mean, sem, expect))
    assert abs(mean - expect) < 4.0 * sem + 1e-3 * expect
    dev.close()

    # sphere light straight above: rho * E * (r/d)^2 (sphere.rs:27-36,104-119)
    E, r, d = 4.0, 1.0, 3.0
    rho, sd = floor_under(lambda b: b.sphere((0.0, d, 0.0), r, b.diffuse_light(b.constant_texture((E, E, E)))))
    dev = rt.DeviceScene(sd, device=0)
    rgb, _ = dev.path_radiance(cam, 3, 3, 50, rt.render_opts(seed=5, integrator=rt.INTEGRATOR_HEAD), px, py, smp)
    expect = rho * E * (r / d) ** 2
    mean, sem = rgb[:, 0].mean(), rgb[:, 0].std() / math.sqrt(n)
    print("sphere light: L = %.5f +- %.5f, closed form %.5f" % (mean, sem, expect))
    assert abs(mean - expect) < 4.0 * sem + 1e-3 * expect
    dev.close()

    # Beer-Lambert through a black-albedo slab (medium.rs:42-45), legacy integrator
    density, thickness = 0.35, 2.0
    b = rt.SceneBuilder()
    slab = b.cube((-50, -50, 0.0), (50, 50, thickness), b.lambertian(b.constant_texture((1, 1, 1))))
    fog = b.medium(slab, density, b.constant_texture((0.0, 0.0, 0.0)))
    sd = b.finish(b.list([fog]), b.list([]), background=(1.0, 1.0, 1.0))
    dev = rt.DeviceScene(sd, device=0)
    cam2 = rt.camera_new((0, 0, -10), (0, 0, 0), (0, 1, 0), 0.05, 1.0, 0.0, 10.0)
    rgb, _ = dev.path_radiance(cam2, 3, 3, 50, rt.render_opts(seed=9, integrator=rt.INTEGRATOR_LEGACY), px, py, smp)
    expect = math.exp(-density * thickness)
    assert set(np.unique(rgb[:, 0])) <= {0.0, 1.0}
    assert abs(rgb[:, 0].mean() - expect) < 4.0 * math.sqrt(expect * (1 - expect) / n)
    dev.close()
