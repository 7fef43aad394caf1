"""Synthetic TRON1 workloads.

Counter-based generator: every scalar is a pure function of (seed, instance, field), so any
consumer (numpy here, the C oracle, a C++ host program) can regenerate the identical batch
without carrying state.  u01(seed, inst, field) = top 53 bits of splitmix64 over a mixed key.

The gait/contact schedule is NOT synthesised here: it is produced by the engine's
contact-schedule entry point (bit-exact restatement of MPC::calculateGait,
reference include/MPCController.h:61-75) from the per-instance `iter` drawn below.
"""
import numpy as np

# constants from the reference (include/MPCParam.h:13-38,64-73)
FOOT_OFFSET_L = np.array([0.05556 - 0.077 - 0.15 + 0.145 + 0.0,
                          -0.105 - 0.0205 - (-0.0205) + 0.0 + 0.0,
                          -0.2602 + 0.0 - 0.25981 - 0.2598 - 0.032])
FOOT_OFFSET_R = np.array([0.05556 - 0.077 - 0.15 + 0.145 + 0.0,
                          0.105 + 0.0205 + (-0.0205) + 0.0 + 0.0,
                          -0.2602 + 0.0 - 0.25981 - 0.2598 - 0.032])

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def u01(seed, inst, field):
    """Uniform [0,1) double keyed by (seed, instance, field); inst may be an array."""
    with np.errstate(over="ignore"):
        inst = np.asarray(inst, dtype=np.uint64)
        k = _splitmix64(np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + np.uint64(field))
        z = _splitmix64(k ^ (inst * np.uint64(0x2545F4914F6CDD1D)))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(seed, inst, field, lo, hi):
    return lo + (hi - lo) * u01(seed, inst, field)


def tron1_batch(seed, B, N, Ts, first=0, per_step_feet=False, standing=False):
    """Returns dict of C-contiguous arrays for instances [first, first+B):
       x0[B,13], x_ref[B,N+1,13], feet[B,2,3] (or [B,N,2,3]), iter[B] int32,
       omega_yaw[B], velocity_x[B].
    Distributions: SURVEY.md 8d.  Reference generator: include/mpcQP.h:74-97 with per-instance
    omega_yaw, velocity_x."""
    inst = np.arange(first, first + B, dtype=np.uint64)
    U = lambda f, lo, hi: uniform(seed, inst, f, lo, hi)
    x0 = np.empty((B, 13))
    x0[:, 0] = U(0, -0.2, 0.2); x0[:, 1] = U(1, -0.2, 0.2); x0[:, 2] = U(2, -np.pi, np.pi)
    x0[:, 3] = U(3, -1, 1); x0[:, 4] = U(4, -1, 1); x0[:, 5] = U(5, 0.70, 0.90)
    x0[:, 6] = U(6, -0.5, 0.5); x0[:, 7] = U(7, -0.5, 0.5); x0[:, 8] = U(8, -0.5, 0.5)
    x0[:, 9] = U(9, -1, 1); x0[:, 10] = U(10, -0.5, 0.5); x0[:, 11] = U(11, -0.2, 0.2)
    x0[:, 12] = -9.8
    omega_yaw = U(12, -0.5, 0.5)
    velocity_x = U(13, -1, 1)
    if standing:
        it = np.full(B, -1, np.int32)   # iter < 0 => both feet in contact for the whole horizon
    else:
        it = np.floor(U(14, 0.0, 1.0e7 + 1.0)).astype(np.int32)
    # reference trajectory (include/mpcQP.h:74-97)
    x_ref = np.repeat(x0[:, None, :], N + 1, axis=1)
    t = np.arange(N + 1) * Ts
    x_ref[:, :, 2] = x0[:, None, 2] + t[None, :] * omega_yaw[:, None]
    x_ref[:, :, 3] = x0[:, None, 3] + t[None, :] * velocity_x[:, None]
    x_ref[:, 1:, 9] = velocity_x[:, None]
    x_ref[:, :, 12] = -9.8
    # feet: p_xy + Rz(yaw) offset_xy + U(-0.1,0.1)^2, z = 0
    c, s = np.cos(x0[:, 2]), np.sin(x0[:, 2])
    feet = np.zeros((B, 2, 3))
    for i, off in enumerate((FOOT_OFFSET_L, FOOT_OFFSET_R)):
        feet[:, i, 0] = x0[:, 3] + c * off[0] - s * off[1] + U(20 + 2 * i, -0.1, 0.1)
        feet[:, i, 1] = x0[:, 4] + s * off[0] + c * off[1] + U(21 + 2 * i, -0.1, 0.1)
    if per_step_feet:
        feet = np.ascontiguousarray(np.repeat(feet[:, None], N, axis=1))
    return dict(x0=np.ascontiguousarray(x0), x_ref=np.ascontiguousarray(x_ref), feet=feet,
                iter=it, omega_yaw=omega_yaw, velocity_x=velocity_x)
