"""CPU restatement of the domain-randomisation noise path (TEST INFRASTRUCTURE: imported by tests/, smoke() and bench.py's
CPU arms only -- never by the product path).

Follows BaseTask.apply_randomizations' non-physical branch (`tasks/hand_base/base_task.py:263-340`: schedule scaling, gaussian /
uniform parameters, the noise_lambda closures) and the two hooks of BaseTask.step (`:131-132`, `:149-150`).  Pinned to the
reference's own Python through tests/golden/dr_params.npz (oracle/gen_golden_dr.py executes apply_randomizations and the
closures it creates).  The white noise itself is this repo's Philox stream (csrc/sdx_dr.cuh), restated here in numpy.
"""
import numpy as np

DR_STREAM = 0x44520000


def schedule_scaling(cfg, last_step):
    """BT:269-277"""
    sched_type = cfg["schedule"] if "schedule" in cfg else None
    sched_step = cfg["schedule_steps"] if "schedule" in cfg else None
    if sched_type == "linear":
        return 1.0 / sched_step * min(last_step, sched_step)
    if sched_type == "constant":
        return 0 if last_step < sched_step else 1
    return 1


def nonphysical_params(cfg, last_step):
    """BT:279-334: the four numbers a noise_lambda closes over, as (distribution, operation, a_corr, b_corr, a, b) in the form
    sdx_dr_noise takes (gaussian: var_corr, mu_corr, var, mu; uniform: hi_corr - lo_corr, lo_corr, hi - lo, lo)."""
    dist, op_type = cfg["distribution"], cfg["operation"]
    s = schedule_scaling(cfg, last_step)
    if dist == "gaussian":
        mu, var = cfg["range"]
        mu_corr, var_corr = cfg.get("range_correlated", [0., 0.])
        if op_type == "additive":
            mu *= s; var *= s; mu_corr *= s; var_corr *= s
        elif op_type == "scaling":
            var = var * s
            mu = mu * s + 1.0 * (1.0 - s)
            var_corr = var_corr * s
            mu_corr = mu_corr * s + 1.0 * (1.0 - s)
        return {"distribution": 0, "operation": int(op_type == "scaling"), "a_corr": var_corr, "b_corr": mu_corr, "a": var, "b": mu,
                "mu": mu, "var": var, "mu_corr": mu_corr, "var_corr": var_corr}
    if dist == "uniform":
        lo, hi = cfg["range"]
        lo_corr, hi_corr = cfg.get("range_correlated", [0., 0.])
        if op_type == "additive":
            lo *= s; hi *= s; lo_corr *= s; hi_corr *= s
        elif op_type == "scaling":
            lo = lo * s + 1.0 * (1.0 - s)
            hi = hi * s + 1.0 * (1.0 - s)
            lo_corr = lo_corr * s + 1.0 * (1.0 - s)
            hi_corr = hi_corr * s + 1.0 * (1.0 - s)
        return {"distribution": 1, "operation": int(op_type == "scaling"), "a_corr": hi_corr - lo_corr, "b_corr": lo_corr, "a": hi - lo,
                "b": lo, "lo": lo, "hi": hi, "lo_corr": lo_corr, "hi_corr": hi_corr}
    raise ValueError(f"unknown distribution {dist!r}")


def combine(src, corr, white, p):
    """the body of noise_lambda (BT:292-299 / 321-327) in fp32, operation for operation"""
    f = np.float32
    c = corr.astype(f) * f(p["a_corr"]) + f(p["b_corr"])
    noise = (c + white.astype(f) * f(p["a"])) + f(p["b"])
    return src.astype(f) * noise if p["operation"] else src.astype(f) + noise


def philox4x32(seed, c0, c1, c2):
    """Philox4x32-10, key = seed, counter = (c0, c1, c2, 0); c0 may be an array (csrc/sdx_math.cuh philox)"""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    c = [np.asarray(c0, dtype=np.uint64) & mask, np.full_like(np.asarray(c0, dtype=np.uint64), c1),
         np.full_like(np.asarray(c0, dtype=np.uint64), c2), np.zeros_like(np.asarray(c0, dtype=np.uint64))]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        n0 = ((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask
        n3 = p0 & mask
        c = [n0, n1, n2, n3]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c


def white(n, seed, counter, uniform):
    """the kernel's white-noise stream: element i comes from Philox quad i // 4, lane i % 4 (csrc/sdx_dr.cuh dr_white4)"""
    f = np.float32
    nq = (n + 3) // 4
    r = philox4x32(seed, np.arange(nq, dtype=np.uint64), counter, DR_STREAM)
    r = [(x >> np.uint64(8)).astype(np.float32) for x in r]
    inv = f(1.0 / 16777216.0)
    if uniform:
        w = np.stack([x * inv for x in r], axis=1)
    else:
        u0, u1, u2, u3 = (r[0] + f(0.5)) * inv, r[1] * inv, (r[2] + f(0.5)) * inv, r[3] * inv
        ra = np.sqrt(f(-2.0) * np.log(u0), dtype=f)
        rb = np.sqrt(f(-2.0) * np.log(u2), dtype=f)
        tp = f(6.283185307179586)
        w = np.stack([ra * np.cos(tp * u1), ra * np.sin(tp * u1), rb * np.cos(tp * u3), rb * np.sin(tp * u3)], axis=1).astype(f)
    return w.reshape(-1)[:n]


def randn(n, seed, counter):
    return white(n, seed, counter, uniform=False)


def noise(src, corr, p, seed, counter):
    flat = np.ascontiguousarray(src, dtype=np.float32).reshape(-1)
    w = white(flat.size, seed, counter, uniform=bool(p["distribution"]))
    return combine(flat, np.ascontiguousarray(corr, dtype=np.float32).reshape(-1), w, p).reshape(np.shape(src))


def generate_random_samples(cfg, shape, curr_step, rng):
    """isaacgym.gymutil.generate_random_samples (Isaac Gym Preview 4, python/isaacgym/gymutil.py -- a third-party file that is
    NOT under /root/reference; restated from the published package): one sample of a physical parameter's perturbation."""
    lo_hi, dist, op = cfg["range"], cfg["distribution"], cfg["operation"]
    s = schedule_scaling(cfg, curr_step)
    if dist == "gaussian":
        mu, var = lo_hi
        if op == "additive":
            mu *= s; var *= s
        elif op == "scaling":
            var = var * s
            mu = mu * s + 1.0 * (1.0 - s)
        return rng.normal(mu, var, shape)
    lo, hi = lo_hi
    if op == "additive":
        lo *= s; hi *= s
    elif op == "scaling":
        lo = lo * s + 1.0 * (1.0 - s)
        hi = hi * s + 1.0 * (1.0 - s)
    if dist == "loguniform":
        return np.exp(rng.uniform(np.log(lo), np.log(hi), shape))
    if dist == "uniform":
        return rng.uniform(lo, hi, shape)
    raise ValueError(f"unknown distribution {dist!r}")
