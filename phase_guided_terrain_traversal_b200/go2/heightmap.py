"""`create_sensor_matrix` (go2/heightmap.py:25-67): 13x9 yaw-aligned vertical ray grid above `center`.

Reference signature is `(mx, dx, center, yaw, key=None)` with a single env; here `env` stands for the
(model, data) pair and `center [N,3]`, `yaw [N]` are batched. Returns hit points `[N,13,9,3]`.
The same rays are cast inside the fused step/reset kernels (csrc/pgtt_env.cuh:heightscan); this entry
point launches the stand-alone kernel behind `pgtt_heightscan`.
"""


def create_sensor_matrix(env, center, yaw):
    abi = env._abi
    if abi is None:
        raise RuntimeError("create_sensor_matrix needs an env whose handle exists (reset or randomize first)")
    torch = abi.torch
    center = torch.as_tensor(center, dtype=torch.float32, device=abi.torch_device).reshape(abi.N, 3)
    yaw = torch.as_tensor(yaw, dtype=torch.float32, device=abi.torch_device).reshape(-1).expand(abi.N).contiguous()
    return abi.heightscan(center, yaw)
