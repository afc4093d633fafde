"""GO2 model compiler (host, float64).

Replaces what `mujoco.MjModel.from_xml_string` + `mjx.put_model` do for the reference
(`go2/base.py:53-68`): turns the constant table extracted from the MJCF
(`assets/go2_model.json`, written by `tools/extract_go2_model.py` from
`go2/xmls/go2_mjx_feetonly.xml:1-292`, `scene_mjx_feetonly.xml:21`,
`terrain_scene_mjx.xml:20-21`) into the flat arrays the oracle and the CUDA kernels consume,
applies the `Go2Env.__init__` overrides (timestep, Kp, Kd: `go2/base.py:57-62`) and computes
the constants MuJoCo's compiler derives at `qpos0` (`body_subtreemass`, `body_invweight0`,
`dof_invweight0`, `stat.meaninertia`).

Everything here is plain numpy in float64 and deliberately written in the generic
"body tree + 6-D Jacobian" form, i.e. structurally different from both the C oracle
(spatial-vector CRBA/RNE) and the CUDA kernel (leg-specialised arrow matrices), so the three
can cross-check one another (tests/test_model.py, tests/test_oracle_physics.py).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

ASSET_DIR = Path(__file__).resolve().parent / "assets"

NBODY = 14          # world + base + 4 x (hip, thigh, calf)
NQ, NV, NU = 19, 18, 12
NLEG = 4
MAX_BOXES = 100
NSENSORDATA = 49
TASK_TO_SCENE = {"flat_terrain": "flat_terrain", "stairs": "stairs"}   # go2/go2_constants.py:45-52


# --------------------------------------------------------------------------------------
# small quaternion / rotation helpers (w, x, y, z)
# --------------------------------------------------------------------------------------
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def axis_angle_to_quat(axis, angle):
    s, c = np.sin(angle * 0.5), np.cos(angle * 0.5)
    return np.array([c, axis[0] * s, axis[1] * s, axis[2] * s])


def load_spec(path: Path | None = None) -> dict:
    with open(path or (ASSET_DIR / "go2_model.json")) as f:
        return json.load(f)


@dataclass
class Go2Model:
    """Flat, nominal (un-randomised) model. Index conventions follow MuJoCo's compiled model:
    bodies 0..13 = world, base, FL_{hip,thigh,calf}, FR_*, RL_*, RR_*; dofs 0..5 free joint
    (3 world-frame translations, 3 body-frame rotations), 6..17 hinges in body order;
    actuators in MJCF order FR, FL, RR, RL."""
    task: str
    n_boxes: int
    timestep: float
    gravity: np.ndarray
    impratio: float
    iterations: int
    ls_iterations: int
    tolerance: float
    ls_tolerance: float
    max_geom_pairs: int
    max_contact_points: int
    body_parent: np.ndarray
    body_pos: np.ndarray
    body_quat: np.ndarray
    body_ipos: np.ndarray
    body_iquat: np.ndarray
    body_mass: np.ndarray
    body_inertia: np.ndarray
    jnt_body: np.ndarray          # [12] body of hinge j
    jnt_axis: np.ndarray          # [12,3] body frame
    jnt_range: np.ndarray         # [12,2]
    jnt_solref: np.ndarray
    jnt_solimp: np.ndarray
    qpos0: np.ndarray             # [19]
    dof_armature: np.ndarray      # [18]
    dof_damping: np.ndarray       # [18]
    act_dof: np.ndarray           # [12] dof driven by actuator a
    act_gainprm: np.ndarray       # [12,3]
    act_biasprm: np.ndarray       # [12,3]
    act_ctrlrange: np.ndarray     # [12,2]
    act_forcerange: np.ndarray    # [12,2]
    foot_body: np.ndarray         # [4] in geom-id order FL, FR, RL, RR
    foot_geom_id: np.ndarray      # [4] = 20, 32, 44, 56
    foot_pos: np.ndarray          # [3] in calf frame (same for all feet; also the foot site)
    foot_radius: float
    foot_friction: np.ndarray     # [3]
    foot_solref: np.ndarray
    foot_solimp: np.ndarray
    foot_margin: float
    floor_geom_id: int
    floor_friction: np.ndarray
    floor_solref: np.ndarray
    floor_solimp: np.ndarray
    box_geom_id0: int             # 57
    box_body_id0: int             # 14
    box_friction: np.ndarray
    box_rbound: float             # sqrt(3): placeholder size (1,1,1), never re-derived
    box_park: np.ndarray          # [n_boxes,10] placeholder rows (pos, quat, size) of the scene file
    imu_pos: np.ndarray           # [3] in base frame
    home_qpos: np.ndarray         # [19]
    home_ctrl: np.ndarray         # [12]
    sensor_adr: dict
    # derived at qpos0 with NOMINAL parameters (never re-derived after DR)
    body_subtreemass: np.ndarray = field(default=None)
    body_invweight0: np.ndarray = field(default=None)   # [14,2]
    dof_invweight0: np.ndarray = field(default=None)    # [18]
    meaninertia: float = 0.0

    # actuator a -> hinge index (0..11 in qpos order) and inverse
    @property
    def act_hinge(self):
        return self.act_dof - 6


# --------------------------------------------------------------------------------------
# generic numpy kinematics / dynamics (float64) - used for derived constants and as an
# independent cross-check of the oracle
# --------------------------------------------------------------------------------------
def kinematics(m: Go2Model, qpos, qpos0=None, body_ipos=None):
    """Returns dict with xpos[14,3], xquat[14,4], xmat[14,3,3], xipos[14,3], ximat[14,3,3],
    xanchor[12,3], xaxis[12,3] (world frame)."""
    qpos = np.asarray(qpos, dtype=np.float64)
    qpos0 = m.qpos0 if qpos0 is None else qpos0
    body_ipos = m.body_ipos if body_ipos is None else body_ipos
    xpos = np.zeros((NBODY, 3))
    xquat = np.zeros((NBODY, 4))
    xquat[0] = [1, 0, 0, 0]
    xanchor = np.zeros((12, 3))
    xaxis = np.zeros((12, 3))
    hinge_of_body = {int(b): j for j, b in enumerate(m.jnt_body)}
    for b in range(1, NBODY):
        p = m.body_parent[b]
        if b == 1:  # free joint: pose straight from qpos (quat normalised)
            xpos[b] = qpos[0:3]
            q = qpos[3:7]
            xquat[b] = q / np.linalg.norm(q)
            continue
        pmat = quat_to_mat(xquat[p])
        pos = xpos[p] + pmat @ m.body_pos[b]
        quat = quat_mul(xquat[p], m.body_quat[b])
        j = hinge_of_body[b]
        xanchor[j] = pos                      # jnt_pos = 0 for every GO2 joint
        xaxis[j] = quat_to_mat(quat) @ m.jnt_axis[j]
        quat = quat_mul(quat, axis_angle_to_quat(m.jnt_axis[j], qpos[7 + j] - qpos0[7 + j]))
        xpos[b] = pos
        xquat[b] = quat / np.linalg.norm(quat)
    xmat = np.stack([quat_to_mat(q) for q in xquat])
    xipos = xpos + np.einsum("bij,bj->bi", xmat, body_ipos)
    ximat = np.stack([xmat[b] @ quat_to_mat(m.body_iquat[b]) for b in range(NBODY)])
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, ximat=ximat,
                xanchor=xanchor, xaxis=xaxis)


def body_chain_dofs(m: Go2Model, b: int):
    """dofs that move body b (free-joint dofs + hinges up the chain)."""
    dofs = []
    hinge_of_body = {int(bb): j for j, bb in enumerate(m.jnt_body)}
    while b > 1:
        dofs.append(6 + hinge_of_body[b])
        b = int(m.body_parent[b])
    if b == 1:
        dofs += [0, 1, 2, 3, 4, 5]
    return sorted(dofs)


def jacobian(m: Go2Model, kin, point, body):
    """6 x nv Jacobian [jacp; jacr] of a world `point` fixed to `body` (MuJoCo mj_jac):
    free joint = 3 world translations + 3 rotations about the BODY axes of the base."""
    jacp = np.zeros((3, NV))
    jacr = np.zeros((3, NV))
    hinge_of_body = {int(bb): j for j, bb in enumerate(m.jnt_body)}
    b = body
    while b > 1:
        j = hinge_of_body[b]
        ax = kin["xaxis"][j]
        jacr[:, 6 + j] = ax
        jacp[:, 6 + j] = np.cross(ax, point - kin["xanchor"][j])
        b = int(m.body_parent[b])
    if b == 1:
        jacp[:, 0:3] = np.eye(3)
        R = kin["xmat"][1]
        for k in range(3):
            jacr[:, 3 + k] = R[:, k]
            jacp[:, 3 + k] = np.cross(R[:, k], point - kin["xpos"][1])
    return jacp, jacr


def mass_matrix(m: Go2Model, qpos, body_mass=None, body_ipos=None, armature=None, qpos0=None):
    """Dense joint-space inertia by summing J^T [m, I] J over bodies (NOT CRBA on purpose)."""
    body_mass = m.body_mass if body_mass is None else body_mass
    armature = m.dof_armature if armature is None else armature
    kin = kinematics(m, qpos, qpos0=qpos0, body_ipos=body_ipos)
    M = np.zeros((NV, NV))
    for b in range(1, NBODY):
        jp, jr = jacobian(m, kin, kin["xipos"][b], b)
        Iw = kin["ximat"][b] @ np.diag(m.body_inertia[b]) @ kin["ximat"][b].T
        M += body_mass[b] * jp.T @ jp + jr.T @ Iw @ jr
    M += np.diag(armature)
    return M, kin


def potential_energy(m: Go2Model, qpos, body_mass=None, body_ipos=None, qpos0=None):
    body_mass = m.body_mass if body_mass is None else body_mass
    kin = kinematics(m, qpos, qpos0=qpos0, body_ipos=body_ipos)
    return float(-(body_mass[:, None] * kin["xipos"] * m.gravity[None, :]).sum())


def _derive_constants(m: Go2Model):
    m.body_subtreemass = np.zeros(NBODY)
    for b in range(NBODY - 1, 0, -1):
        m.body_subtreemass[b] += m.body_mass[b]
        m.body_subtreemass[m.body_parent[b]] += m.body_subtreemass[b]
    M, kin = mass_matrix(m, m.qpos0)
    Minv = np.linalg.inv(M)
    m.meaninertia = float(np.mean(np.diag(M)))
    dinv = np.diag(Minv).copy()
    dinv[0:3] = dinv[0:3].mean()       # free joint: averaged per translational / rotational triple
    dinv[3:6] = dinv[3:6].mean()
    m.dof_invweight0 = dinv
    m.body_invweight0 = np.zeros((NBODY, 2))
    for b in range(1, NBODY):
        jp, jr = jacobian(m, kin, kin["xipos"][b], b)
        J = np.vstack([jp, jr])
        A = J @ Minv @ J.T
        m.body_invweight0[b, 0] = (A[0, 0] + A[1, 1] + A[2, 2]) / 3.0
        m.body_invweight0[b, 1] = (A[3, 3] + A[4, 4] + A[5, 5]) / 3.0


def compile_model(task: str = "flat_terrain", sim_dt: float = 0.005, Kp: float = 40.0,
                  Kd: float = 0.5, spec: dict | None = None) -> Go2Model:
    """`Go2Env.__init__` for the constants (go2/base.py:45-113). Raises KeyError for a task the
    reference's `task_to_xml` does not know or whose scene file does not exist."""
    spec = spec or load_spec()
    scene = spec["scenes"][TASK_TO_SCENE[task]]
    bodies = spec["bodies"]
    assert len(bodies) == NBODY
    name_to_body = {b["name"]: i for i, b in enumerate(bodies)}
    hinges = [j for j in spec["joints"] if j["type"] == "hinge"]
    assert len(hinges) == 12 and spec["joints"][0]["type"] == "free"
    jnt_name_to_idx = {j["name"]: i for i, j in enumerate(hinges)}
    opt = spec["option"]
    assert opt["cone"] == "pyramidal" and opt["integrator"] == "Euler" and not opt["eulerdamp"]

    qpos0 = np.zeros(NQ)
    qpos0[0:3] = bodies[1]["pos"]
    qpos0[3:7] = bodies[1]["quat"]
    feet = sorted(spec["geoms"], key=lambda g: g["id"])
    assert [g["name"] for g in feet] == ["FL", "FR", "RL", "RR"]
    site_pos = {s["name"]: np.array(s["pos"]) for s in spec["sites"]}
    for g in feet:  # foot site coincides with the foot sphere centre
        assert np.allclose(site_pos[g["name"] + "_foot"], g["pos"])
    nb = scene["n_boxes"]
    tpl = scene["box_template"]
    park = np.zeros((nb, 10))
    if nb:
        for k in range(nb):
            park[k, 0:3] = np.array(tpl["pos0"]) + k * np.array(tpl["pos_step"])
            park[k, 3:7] = tpl["quat"]
            park[k, 7:10] = tpl["size"]

    damping = np.zeros(NV)
    damping[6:] = Kd                                              # base.py:60
    act_gain = np.array([a["gainprm"] for a in spec["actuators"]], dtype=np.float64)
    act_bias = np.array([a["biasprm"] for a in spec["actuators"]], dtype=np.float64)
    act_gain[:, 0] = Kp                                           # base.py:61
    act_bias[:, 1] = -Kp                                          # base.py:62

    m = Go2Model(
        task=task, n_boxes=nb, timestep=float(sim_dt),            # base.py:57
        gravity=np.array(opt["gravity"]), impratio=opt["impratio"],
        iterations=opt["iterations"], ls_iterations=opt["ls_iterations"],
        tolerance=opt["tolerance"], ls_tolerance=opt["ls_tolerance"],
        max_geom_pairs=int(spec["numeric"]["max_geom_pairs"]),
        max_contact_points=int(spec["numeric"]["max_contact_points"]),
        body_parent=np.array([max(b["parent"], 0) for b in bodies]),
        body_pos=np.array([b["pos"] for b in bodies], dtype=np.float64),
        body_quat=np.array([b["quat"] for b in bodies], dtype=np.float64),
        body_ipos=np.array([b["ipos"] for b in bodies], dtype=np.float64),
        body_iquat=np.array([np.array(b["iquat"]) / np.linalg.norm(b["iquat"]) for b in bodies]),
        body_mass=np.array([b["mass"] for b in bodies], dtype=np.float64),
        body_inertia=np.array([b["inertia"] for b in bodies], dtype=np.float64),
        jnt_body=np.array([j["body"] for j in hinges]),
        jnt_axis=np.array([j["axis"] for j in hinges], dtype=np.float64),
        jnt_range=np.array([j["range"] for j in hinges], dtype=np.float64),
        jnt_solref=np.array(hinges[0]["solref_limit"]), jnt_solimp=np.array(hinges[0]["solimp_limit"]),
        qpos0=qpos0,
        dof_armature=np.concatenate([np.zeros(6), [j["armature"] for j in hinges]]),
        dof_damping=damping,
        act_dof=np.array([6 + jnt_name_to_idx[a["joint"]] for a in spec["actuators"]]),
        act_gainprm=act_gain, act_biasprm=act_bias,
        act_ctrlrange=np.array([a["ctrlrange"] for a in spec["actuators"]]),
        act_forcerange=np.array([a["forcerange"] for a in spec["actuators"]]),
        foot_body=np.array([g["body"] for g in feet]),
        foot_geom_id=np.array([g["id"] for g in feet]),
        foot_pos=np.array(feet[0]["pos"]), foot_radius=feet[0]["size"][0],
        foot_friction=np.array(feet[0]["friction"]), foot_solref=np.array(feet[0]["solref"]),
        foot_solimp=np.array(feet[0]["solimp"]), foot_margin=feet[0]["margin"],
        floor_geom_id=0, floor_friction=np.array(scene["floor"]["friction"]),
        floor_solref=np.array(scene["floor"]["solref"]), floor_solimp=np.array(scene["floor"]["solimp"]),
        box_geom_id0=spec["n_robot_geoms_end"], box_body_id0=NBODY,
        box_friction=np.array((tpl or scene["floor"])["friction"]),
        box_rbound=float(np.linalg.norm(tpl["size"])) if tpl else 0.0,
        box_park=park,
        imu_pos=site_pos["imu"],
        home_qpos=np.array(spec["keyframe_home"]["qpos"]),
        home_ctrl=np.array(spec["keyframe_home"]["ctrl"]),
        sensor_adr={s["name"]: (s["adr"], s["dim"]) for s in spec["sensors"]},
    )
    for j in hinges:
        assert j["solref_limit"] == hinges[0]["solref_limit"] and j["pos"] == [0, 0, 0]
    assert name_to_body["base"] == 1 and list(m.foot_body) == [4, 7, 10, 13]
    _derive_constants(m)
    return m
