"""pgtt-b200: B200-native batched GO2 joystick environment (env.reset / env.step, rollout collector, PPO learner) behind the
MuJoCo-Playground-style API of the PGTT training script. Python host over the C ABI `include/pgtt_b200.h`
(`csrc/libpgtt_b200.so`, built by `python __graft_entry__.py`); see DESIGN.md and INTEGRATION.md."""

__version__ = "0.1.0"
