"""Host-side mirror of the reference's `go2` package for the hot path: `configs`, `go2_constants`, `base.Go2Env` / `State`,
`joystick_pgtt.Joystick` (PGTT task), `joystick.Joystick` (baseline task), `randomize`, `randomize_simple`."""
