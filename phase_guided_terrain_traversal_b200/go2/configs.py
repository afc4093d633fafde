"""Env configuration factories - same keys and values as go2/configs.py:6-152."""
from .. import config_dict
from . import go2_constants as consts


def _common(reward_scales, tracking_sigma):
    return config_dict.create(
        ctrl_dt=0.02,
        sim_dt=0.005,
        episode_length=1000,
        vel_percentage=0.65,
        Kp=40.0,
        Kd=0.5,
        action_repeat=1,
        action_scale=0.5,
        history_len=2,
        history_update_steps=5,
        soft_joint_pos_limit_factor=0.95,
        noise_config=config_dict.create(
            level=1.0,  # 0.0 disables observation noise
            scales=config_dict.create(
                joint_pos=0.03, joint_vel=1.5, gyro=0.2, gravity=0.05, linvel=0.1, heightscan=0.01,
            ),
        ),
        reward_config=config_dict.create(
            scales=config_dict.create(**reward_scales),
            tracking_sigma=tracking_sigma,
            swing_height=-0.2,
            base_feet_distance=-0.3,
            phase_sigma=0.05,
        ),
        command_config=config_dict.create(
            u_max=[1.5, 0.8, 1.2],
            u_min=[-1.5, -0.8, -1.2],
            b=[0.9, 0.25, 0.5],  # probability of NOT zeroing a freshly drawn command component
        ),
        gait_freq=[2, 6],
        heighmap_size=(consts.num_heightscans, consts.num_widthscans),
    )


# key order == metrics order == go2/configs.py:31-59
_PGTT_SCALES = dict(
    tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-1.0, ang_vel_xy=-0.05, orientation=-0.2,
    dof_pos_limits=-1.0, pose=-1.0, termination=-1.0, stand_still=-0.0, torques=-0.0002,
    action_rate=-0.01, energy=-0.0005, feet_clearance=-0.0, feet_height=-0.0, feet_slip=-0.0,
    feet_air_time=0.0, feet_phase=0.5, feet_swing=0.0, body_height=-0.0, contact=2.0, center=-0.0,
)
_BASELINE_SCALES = dict(
    tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-2.0, ang_vel_xy=-0.05, orientation=-0.2,
    dof_pos_limits=-1.0, pose=-0.2, termination=-1.0, stand_still=-0.5, torques=-0.0002,
    action_rate=-0.005, energy=-0.0005, feet_clearance=-1.0, feet_height=-0.0, feet_slip=-0.1,
    feet_air_time=0.1, feet_phase=0.0, feet_swing=0.0, body_height=-0.0, contact=0.0, center=-0.0,
)


def default_config() -> config_dict.ConfigDict:
    """PGTT task config (go2/configs.py:6-79)."""
    return _common(_PGTT_SCALES, tracking_sigma=0.2)


def baseline_config() -> config_dict.ConfigDict:
    """Non-phase baseline config (go2/configs.py:82-152)."""
    return _common(_BASELINE_SCALES, tracking_sigma=0.25)


def training_overrides(cfg: config_dict.ConfigDict) -> config_dict.ConfigDict:
    """The mutations training/train.py:127-129 applies after building the config."""
    cfg.command_config.u_max = [0.6, 0.6, 1.0]
    cfg.command_config.u_min = [-0.6, -0.6, -1.0]
    cfg.gait_freq = [1, 3]
    return cfg


def eval_overrides(cfg: config_dict.ConfigDict) -> config_dict.ConfigDict:
    """The mutations training/evaluate.py:127-129 applies: narrower command ranges than training."""
    cfg.command_config.u_max = [0.4, 0.4, 0.7]
    cfg.command_config.u_min = [-0.4, -0.4, -0.7]
    cfg.gait_freq = [1, 3]
    return cfg
