"""GO2 constants: names, ids and the height-scan grid (mirrors go2/go2_constants.py:21-94).

The reference maps a task name to an MJCF path; here the MJCF has already been compiled into
`assets/go2_model.json`, so `task_to_xml` returns the scene key understood by `model.compile_model`
and raises the same `KeyError` for an unknown task (go2_constants.py:45-52).
"""
from pathlib import Path

ROOT_PATH = Path("go2")
_TASKS = {
    "flat_terrain": ROOT_PATH / "xmls" / "scene_mjx_feetonly.xml",
    "stairs": ROOT_PATH / "xmls" / "terrain_scene_mjx.xml",
}


def task_to_xml(task_name: str) -> Path:
    return _TASKS[task_name]


FEET_SITES = ["FR_foot", "FL_foot", "RR_foot", "RL_foot"]
FEET_GEOMS = ["FR", "FL", "RR", "RL"]
FEET_POS_SENSOR = ["FR_pos", "FL_pos", "RR_pos", "RL_pos"]
ROOT_BODY = "base"
UPVECTOR_SENSOR = "upvector"
GLOBAL_LINVEL_SENSOR = "global_linvel"
GLOBAL_ANGVEL_SENSOR = "global_angvel"
LOCAL_LINVEL_SENSOR = "local_linvel"
ACCELEROMETER_SENSOR = "accelerometer"
GYRO_SENSOR = "gyro"

# height-scan grid (go2_constants.py:90-94): 13 x 9 rays, 0.1 m pitch
num_heightscans = 13
num_widthscans = 9
dist_x = 0.1
dist_y = 0.1
