"""CPU: the plain-C oracle restates the same reference lines as the NumPy oracle; they must
agree to rounding on every leaf (this pins the C one, which then serves full-size parity)."""

import time

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import jaxsim_oracle as O

from . import helpers as H

MODELS = ["pendulum", "double_pendulum", "cartpole", "box", "sphere", "icub_like", "ergocub_like"]


@pytest.mark.parametrize("in_contact", [False, True])
@pytest.mark.parametrize("name", MODELS)
def test_c_oracle_matches_numpy_oracle(name, in_contact):
    om = H.oracle_model(H.build_model(name))
    B = 9
    od = O.random_model_data(om, B, seed=3, in_contact=in_contact)
    rng = np.random.default_rng(0)
    tau = 10 * rng.uniform(size=(B, om.dofs()))
    W_f = rng.uniform(-1, 1, size=(B, om.number_of_links(), 6))
    if not om.floating_base:
        W_f[:, 0] = 0
    od.tangential_deformation = 1e-4 * rng.uniform(-1, 1, size=od.tangential_deformation.shape)
    ref = O.step(om, od, link_forces_inertial=W_f, joint_force_references=tau)
    got = CO.step(om, od, link_forces_inertial=W_f, joint_force_references=tau, nthreads=2)
    for name_, _ in H.LEAVES:
        a, b = getattr(got, name_), getattr(ref, name_)
        if b.size:
            assert H.rel_err(a, b) <= 1e-10, name_
    if ref.tangential_deformation.size:
        assert np.abs(got.tangential_deformation - ref.tangential_deformation).max() <= 1e-15 + 1e-10 * np.abs(ref.tangential_deformation).max()


def test_c_oracle_full_size_batch_is_fast():
    om = H.oracle_model(H.build_model("icub_like"))
    od = O.random_model_data(om, 4096, seed=1, in_contact=True)
    t0 = time.perf_counter()
    out = CO.step(om, od)
    dt = time.perf_counter() - t0
    assert np.all(np.isfinite(out.link_transforms)) and dt < 30.0
