import numpy as np


def lattice_bank(scene, per_type, seed=0, jitter=0.01):
    """cheap heap bank for parity tests: the initial lattice (GS:737-742) with xy jitter."""
    rng = np.random.default_rng(seed)
    rows = np.ctypeslib.as_array(scene.c.brick_init).reshape(72, 13).astype(np.float32)
    bank = np.broadcast_to(rows, (8, per_type, 72, 13)).copy()
    bank[..., 0:2] += rng.uniform(-jitter, jitter, size=bank[..., 0:2].shape).astype(np.float32)
    bank[..., 7:13] = 0
    return bank
