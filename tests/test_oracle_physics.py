"""Pins the CPU oracle's rigid-body dynamics to physics itself (the reference ships no golden
vectors): independent numpy formulations and conservation laws."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from phase_guided_terrain_traversal_b200 import model as gm


def rand_state(rng, z=1.0):
    qpos = np.zeros(19)
    qpos[0:3] = [rng.uniform(-1, 1), rng.uniform(-1, 1), z]
    q = rng.normal(size=4)
    qpos[3:7] = q / np.linalg.norm(q)
    qpos[7:] = np.array([0, 0.9, -1.8] * 4) + rng.uniform(-0.4, 0.4, 12)
    qvel = rng.normal(size=18) * np.array([1, 1, 1, 2, 2, 2] + [4] * 12)
    return qpos, qvel


def test_mass_matrix_matches_jacobian_sum(flat_model, train_cfg):
    rng = np.random.default_rng(0)
    orc = Oracle(flat_model, train_cfg, 4, "f64")
    qs = [rand_state(rng) for _ in range(4)]
    orc.set("qpos", np.stack([q for q, _ in qs]))
    orc.set("qvel", np.zeros((4, 18)))
    orc.set("ctrl", np.stack([q[7:] for q, _ in qs]))
    orc.forward()
    M = orc.get("qM").reshape(4, 18, 18)
    for i, (q, _) in enumerate(qs):
        Mref, kin = gm.mass_matrix(flat_model, q)
        assert np.allclose(M[i], Mref, rtol=1e-10, atol=1e-12)
        assert np.allclose(M[i], M[i].T)
        assert np.linalg.eigvalsh(M[i]).min() > 0
        assert np.allclose(orc.get("xpos")[i].reshape(14, 3), kin["xpos"], atol=1e-12)
        assert np.allclose(orc.get("ximat")[i].reshape(14, 3, 3)[1:], kin["ximat"][1:], atol=1e-12)


def test_gravity_bias_is_potential_gradient(flat_model, train_cfg):
    rng = np.random.default_rng(1)
    orc = Oracle(flat_model, train_cfg, 1, "f64")
    q, _ = rand_state(rng)
    orc.set("qpos", q); orc.set("qvel", np.zeros(18)); orc.set("ctrl", q[7:])
    orc.forward()
    bias = orc.get("qfrc_bias")[0]
    eps = 1e-6
    for j in range(12):
        qp, qm = q.copy(), q.copy()
        qp[7 + j] += eps; qm[7 + j] -= eps
        dV = (gm.potential_energy(flat_model, qp) - gm.potential_energy(flat_model, qm)) / (2 * eps)
        assert abs(bias[6 + j] - dV) < 1e-6
    assert abs(bias[2] - flat_model.body_mass.sum() * 9.81) < 1e-9
    assert np.allclose(bias[0:2], 0, atol=1e-12)


def test_free_fall_and_actuator(flat_model, train_cfg):
    orc = Oracle(flat_model, train_cfg, 1, "f64")
    q = flat_model.home_qpos.copy(); q[2] = 2.0
    orc.set("qpos", q); orc.set("qvel", np.zeros(18)); orc.set("ctrl", q[7:])
    orc.forward()
    qacc = orc.get("qacc")[0]
    # PD torque is zero at ctrl == q, so every body falls with g: joint accelerations vanish
    assert np.allclose(qacc[2], -9.81, atol=1e-9)
    assert np.allclose(np.delete(qacc, 2), 0, atol=1e-8)
    # accelerometer reads 0 in free fall, +g at rest is checked in test_standing
    assert np.allclose(orc.get("sensordata")[0, 3:6], 0, atol=1e-8)
    # actuator: kp (ctrl - q) - kv qd, clamped at +-24, FR FL RR RL order
    ctrl = q[7:].copy(); ctrl[1] += 0.1          # actuator 1 = FR_thigh -> dof 10
    orc.set("ctrl", ctrl); orc.forward()
    f = orc.get("actuator_force")[0]
    assert np.isclose(f[1], 40 * 0.1) and np.count_nonzero(f) == 1
    assert np.isclose(orc.get("qfrc_actuator")[0, 10], 4.0)
    ctrl[1] = 2.4; orc.set("ctrl", ctrl); orc.forward()
    assert orc.get("actuator_force")[0, 1] == 24.0
    ctrl[1] = 5.0; orc.set("ctrl", ctrl); orc.forward()   # ctrlrange upper 2.5 -> 40*(2.5-.9) = 64 -> clamp 24
    assert orc.get("actuator_force")[0, 1] == 24.0


def _energy_momentum(model, q, v):
    M, kin = gm.mass_matrix(model, q)
    T = 0.5 * v @ M @ v
    V = gm.potential_energy(model, q)
    # linear / angular momentum about the world origin from body twists
    p = np.zeros(3); L = np.zeros(3)
    for b in range(1, 14):
        jp, jr = gm.jacobian(model, kin, kin["xipos"][b], b)
        vb, wb = jp @ v, jr @ v
        Iw = kin["ximat"][b] @ np.diag(model.body_inertia[b]) @ kin["ximat"][b].T
        p += model.body_mass[b] * vb
        L += np.cross(kin["xipos"][b], model.body_mass[b] * vb) + Iw @ wb
    com = (model.body_mass[:, None] * kin["xipos"]).sum(0) / model.body_mass.sum()
    return T + V, p, L - np.cross(com, p)


def test_free_flight_conserves_energy_and_momentum(flat_model, train_cfg):
    """No damping, no actuation, no contact: E, horizontal momentum and spin about the COM are
    invariants; this exercises kinematics, CRBA, RNE (Coriolis/centrifugal) and the integrator."""
    import copy
    m = copy.deepcopy(flat_model)
    m.dof_damping[:] = 0; m.act_gainprm[:] = 0; m.act_biasprm[:] = 0
    m.dof_armature[:] = 0                       # armature is not part of the rigid-body energy
    m.timestep = 2e-4
    orc = Oracle(m, train_cfg, 1, "f64")
    rng = np.random.default_rng(2)
    q, v = rand_state(rng, z=50.0)
    orc.set("qpos", q); orc.set("qvel", v); orc.set("ctrl", q[7:])
    E0, p0, L0 = _energy_momentum(m, q, v)
    v[6:] *= 0.5                                  # stay clear of the joint limits (they dissipate)
    orc.set("qvel", v)
    E0, p0, L0 = _energy_momentum(m, q, v)
    T0 = E0 - gm.potential_energy(m, q)
    nstep = 250
    for _ in range(nstep):
        orc.physics_step()
        assert np.abs(orc.get("efc_force")).max() == 0.0
    q1, v1 = orc.get("qpos")[0], orc.get("qvel")[0]
    E1, p1, L1 = _energy_momentum(m, q1, v1)
    t = nstep * m.timestep
    # first-order integrator: errors are O(dt); bounds are ~5x the measured drift at this dt
    assert abs(E1 - E0) < 2e-3 * T0
    assert np.allclose(p1[:2], p0[:2], atol=2e-3)
    assert np.isclose(p1[2], p0[2] - m.body_mass.sum() * 9.81 * t, atol=2e-3)
    assert np.allclose(L1, L0, atol=2e-3 * np.linalg.norm(L0))


def test_standing_on_floor(flat_model, train_cfg):
    orc = Oracle(flat_model, train_cfg, 1, "f64")
    q = flat_model.home_qpos.copy()
    orc.set("qpos", q); orc.set("qvel", np.zeros(18)); orc.set("ctrl", q[7:])
    for _ in range(400):                          # 2 s
        orc.physics_step()
    orc.forward()
    q1 = orc.get("qpos")[0]
    assert 0.2 < q1[2] < 0.32 and np.abs(orc.get("qvel")[0]).max() < 0.05
    f, k = orc.contacts(0)
    assert (f[:4, 0] < 0).all() and (k[:4, 4] == -1).all()
    # total normal force = weight: sum of efc_force over pyramid rows (each row ~ normal + mu*tangent)
    ef = orc.get("efc_force")[0]
    J = orc.get("efc_J")[0].reshape(44, 18)
    fz = (J.T @ ef)[2]
    assert np.isclose(fz, flat_model.body_mass.sum() * 9.81, rtol=2e-2)
    s = orc.get("sensordata")[0]
    assert np.allclose(s[3:6], [0, 0, 9.81], atol=0.3)     # accelerometer at rest
    assert s[24] > 0.99                                     # up-vector


def _platform_table(h):
    """One terrain whose box 0 is a 6 m x 6 m platform of height h under the robot; the other 99 boxes are pebbles 50 m away."""
    table = np.zeros((1, 100, 10), np.float32)
    table[0, :, 3] = 1.0
    table[0, :, 0] = 50 + np.arange(100); table[0, :, 1] = 50; table[0, :, 2] = 0.01; table[0, :, 7:] = 0.01
    table[0, 0] = [0, 0, h / 2, 1, 0, 0, 0, 3, 3, h / 2]
    return table


def test_box_contacts_reproduce_plane_contacts_on_a_platform(train_cfg):
    """Translation invariance as a cross-check of two independent code paths: a robot on a large box of height h (sphere/box
    collision, box contact slots, culling, box friction) must move exactly like a robot on the floor (sphere/plane path),
    shifted by h. 200 physics steps with random joint targets, float64 oracle: agreement to 1e-7."""
    h = 0.2
    mf, ms = gm.compile_model("flat_terrain"), gm.compile_model("stairs")
    keys = np.array([[0, 1]], np.uint32)
    of, ob = Oracle(mf, train_cfg, 1, "f64"), Oracle(ms, train_cfg, 1, "f64")
    of.randomize(keys, None, False); ob.randomize(keys, _platform_table(h), False)
    q = mf.home_qpos.copy()
    for o, dz in ((of, 0.0), (ob, h)):
        qq = q.copy(); qq[2] += dz
        o.set("qpos", qq[None]); o.set("qvel", np.zeros((1, 18))); o.set("ctrl", q[7:][None])
    rng = np.random.default_rng(0)
    for s in range(200):
        c = q[7:] + 0.2 * rng.uniform(-1, 1, 12)
        of.set("ctrl", c[None]); ob.set("ctrl", c[None])
        of.physics_step(); ob.physics_step()
    a, b = of.get("qpos")[0], ob.get("qpos")[0].copy()
    b[2] -= h
    assert np.abs(a - b).max() < 1e-7 and np.abs(of.get("qvel")[0] - ob.get("qvel")[0]).max() < 1e-6
    _, kf = of.contacts(0)
    fb, kb = ob.contacts(0)
    assert (kf[:4, 4] == -1).all() and (kf[4:, 0] == -1).all()              # floor run: four plane slots, no box slot
    assert (kb[4:, 4] == 0).all() and (fb[4:, 0] < 0).all()                 # platform run: all four feet penetrate box 0 ...
    assert (fb[:4, 0] > 0.19).all()                                         # ... and are 0.2 m above the plane


def _qmul(a, b):
    w1, x1, y1, z1 = a; w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def quarter_turn(qpos, qvel):
    """The state rotated by +90 degrees about the world z axis (free-joint angular velocity and joints are body-local)."""
    q, v = qpos.copy(), qvel.copy()
    q[0], q[1] = -qpos[1], qpos[0]
    q[3:7] = _qmul(np.array([np.sqrt(0.5), 0, 0, np.sqrt(0.5)]), qpos[3:7])
    v[0], v[1] = -qvel[1], qvel[0]
    return q, v


def mirror_legs(x):
    """Joint (FL FR RL RR) or actuator (FR FL RR RL) triples under the left/right mirror: sides swap, abduction changes sign."""
    y = np.array(x, dtype=np.float64).reshape(4, 3)[[1, 0, 3, 2]]
    y[:, 0] *= -1
    return y.reshape(-1)


def mirror_state(qpos, qvel):
    """Reflection in the x-z plane: positions and velocities flip y, rotations and angular velocities (pseudo-vectors) flip x and z."""
    q, v = qpos.copy(), qvel.copy()
    q[1] *= -1; q[4] *= -1; q[6] *= -1
    q[7:] = mirror_legs(qpos[7:])
    v[1] *= -1; v[3] *= -1; v[5] *= -1
    v[6:] = mirror_legs(qvel[6:])
    return q, v


def _moving_state(m, seed):
    rng = np.random.default_rng(seed)
    q = m.home_qpos.copy(); q[7:] += rng.uniform(-0.2, 0.2, 12); q[2] = 0.3
    qq = np.array([1.0, 0, 0, 0]) + rng.normal(size=4) * 0.05
    q[3:7] = qq / np.linalg.norm(qq)
    v = rng.normal(size=18) * np.array([.3] * 3 + [.5] * 3 + [1] * 12)
    return q, v, m.home_qpos[7:][[3, 4, 5, 0, 1, 2, 9, 10, 11, 6, 7, 8]] + rng.uniform(-0.2, 0.2, 12)


@pytest.mark.parametrize("seed", [3, 8])
def test_quarter_turn_and_mirror_symmetry(flat_model, train_cfg, seed):
    """Symmetries of the robot-on-a-plane system that the implementation does not know about:
    * a quarter turn about z maps trajectories onto trajectories EXACTLY (the friction pyramid is aligned with the world
      axes, so only multiples of 90 degrees are symmetries; 0.7 rad gives 4e-3 after 150 steps) - pins quaternion, body-local
      angular velocity and contact-frame conventions;
    * the left/right mirror does so up to the trunk's slightly asymmetric inertia (2e-5 after 150 steps) - pins the per-leg
      constant tables (axes, offsets, inertia signs) against each other."""
    m = flat_model
    q0, v0, ctrl = _moving_state(m, seed)
    runs = []
    for (q, v), c in (((q0, v0), ctrl), (quarter_turn(q0, v0), ctrl), (mirror_state(q0, v0), mirror_legs(ctrl))):
        o = Oracle(m, train_cfg, 1, "f64")
        o.set("qpos", q[None]); o.set("qvel", v[None]); o.set("ctrl", c[None])
        for _ in range(150):
            o.physics_step()
        runs.append((o.get("qpos")[0].copy(), o.get("qvel")[0].copy()))
    (qa, va), (qt, vt), (qm, vm) = runs
    qe, ve = quarter_turn(qa, va)
    assert np.abs(qe - qt).max() < 1e-10 and np.abs(ve - vt).max() < 1e-9
    qe, ve = mirror_state(qa, va)
    assert np.abs(qe - qm).max() < 2e-4 and np.abs(ve - vm).max() < 2e-3
    assert np.abs(qa[7:] - mirror_legs(qa[7:])).max() > 1e-2            # the motion itself is not symmetric
