"""rl_games ``.pth`` checkpoints <-> the flat fp32 parameter vectors of the CUDA MLPs.

What the reference reads (so what a file we write must hold, and what a file it wrote gives us):
  * ``torch_ext.load_checkpoint(fn)['model']`` -> ``model.load_state_dict`` (utils/robot_controller/nn_controller.py:55-58,
    utils/rl_games_custom.py RGC:2098-2106, ``_restore`` in policy_sequencing/policy_seq_runner.py PSR:74-75,86-87);
  * ``['running_mean_std']`` when ``normalize_input`` (RGC:2101-2102; False in cfg/lego/ppo_continuous_grasp.yaml:64);
  * the agent's full state (RGC:1913-1933 extends rl_games' ``get_full_state_weights``): ``epoch``, ``optimizer``,
    ``assymetric_vf_nets`` (the central-value net, yaml:74-95), ``frame``, ``last_mean_rewards``, ``env_state``.

rl_games itself is a third-party dependency that is absent here (requirements.txt:6 pins 1.5.2); the key names below are
restated from its published network builder (``A2CBuilder.Network``: ``actor_mlp`` / ``critic_mlp`` = ``nn.Sequential`` of
Linear, activation, ...; heads ``mu``, ``value``; parameter ``sigma``) wrapped as ``a2c_network`` by
``ModelA2CContinuousLogStd``.  Everything here is plain tensor bookkeeping on the host: no GPU, no kernels.
"""
from __future__ import annotations

import math
import os

import torch

HIDDEN = (1024, 512, 256)          # cfg/lego/ppo_continuous_grasp.yaml:21-23


def mlp_slices(in_dim, out_dim, hidden=HIDDEN, has_sigma=False):
    """[(name, offset, shape)] of a flat MLP parameter vector, torch state_dict order (ppo.MLP.slices)."""
    dims = [in_dim, *hidden, out_dim]
    out, off = [], 0
    for l in range(len(dims) - 1):
        o, i = dims[l + 1], dims[l]
        out.append((f"W{l}", off, (o, i))); off += o * i
        out.append((f"b{l}", off, (o,))); off += o
    if has_sigma:
        out.append(("sigma", off, (out_dim,))); off += out_dim
    return out, off


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


def default_trunk(in_dim, hidden=HIDDEN, seed=0):
    """torch.nn.Linear-default initialised trunk + 1-d value head (the a2c net's own critic: present in every rl_games
    checkpoint of a ``separate: True`` network, never trained when a central value net exists)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    dims = [in_dim, *hidden]
    for l in range(len(hidden)):
        b = 1.0 / math.sqrt(dims[l])
        sd[f"a2c_network.critic_mlp.{2 * l}.weight"] = (torch.rand(dims[l + 1], dims[l], generator=g) * 2 - 1) * b
        sd[f"a2c_network.critic_mlp.{2 * l}.bias"] = (torch.rand(dims[l + 1], generator=g) * 2 - 1) * b
    b = 1.0 / math.sqrt(hidden[-1])
    sd["a2c_network.value.weight"] = (torch.rand(1, hidden[-1], generator=g) * 2 - 1) * b
    sd["a2c_network.value.bias"] = (torch.rand(1, generator=g) * 2 - 1) * b
    return sd


def actor_state_dict(flat, in_dim, out_dim, hidden=HIDDEN, critic=None, seed=0):
    """flat actor vector (trunk, mu head, sigma) -> ``checkpoint['model']`` of continuous_a2c_logstd / actor_critic,
    separate: True.  ``critic`` = the a2c net's own critic tensors to carry through (from an imported file), else
    default-initialised ones (they carry no information under a central value net)."""
    flat = flat.detach().float().cpu()
    sl, n = mlp_slices(in_dim, out_dim, hidden, has_sigma=True)
    assert flat.numel() == n, f"actor vector has {flat.numel()} parameters, expected {n}"
    sd = {}
    nl = len(hidden)
    for name, off, shp in sl:
        t = flat[off:off + _numel(shp)].view(shp).clone()
        if name == "sigma":
            sd["a2c_network.sigma"] = t
            continue
        l, kind = int(name[1:]), ("weight" if name[0] == "W" else "bias")
        sd[(f"a2c_network.actor_mlp.{2 * l}." if l < nl else "a2c_network.mu.") + kind] = t
    sd.update({k: v.clone() for k, v in (critic or default_trunk(in_dim, hidden, seed)).items()})
    return sd


def _find(sd, suffix):
    """value of the one key that ends with ``suffix`` (tolerates the wrapper prefixes of different rl_games releases:
    ``a2c_network.``, ``model.``, ``model.a2c_network.``)"""
    hits = [k for k in sd if k == suffix or k.endswith("." + suffix)]
    if len(hits) != 1:
        raise KeyError(f"checkpoint has {len(hits)} keys matching '*.{suffix}' (keys: {sorted(sd)[:8]} ...)")
    return sd[hits[0]]


def actor_flat(model_sd, in_dim, out_dim, hidden=HIDDEN):
    """``checkpoint['model']`` -> (flat actor vector, the a2c net's own critic tensors or None)."""
    sl, n = mlp_slices(in_dim, out_dim, hidden, has_sigma=True)
    flat = torch.zeros(n)
    nl = len(hidden)
    for name, off, shp in sl:
        if name == "sigma":
            t = _find(model_sd, "sigma")
        else:
            l, kind = int(name[1:]), ("weight" if name[0] == "W" else "bias")
            t = _find(model_sd, (f"actor_mlp.{2 * l}." if l < nl else "mu.") + kind)
        if tuple(t.shape) != tuple(shp):
            raise ValueError(f"{name}: checkpoint shape {tuple(t.shape)} != network shape {tuple(shp)}")
        flat[off:off + _numel(shp)] = t.detach().float().cpu().reshape(-1)
    critic = {k: v for k, v in model_sd.items() if ".critic_mlp." in k or k.endswith("a2c_network.value.weight") or k.endswith("a2c_network.value.bias")}
    return flat, (critic or None)


def central_value_state_dict(flat, state_dim, hidden=HIDDEN, rms=None):
    """flat central-value vector -> ``checkpoint['assymetric_vf_nets']`` (``CentralValueTrain.state_dict()``: the network
    under ``model.``; its input RunningMeanStd -- yaml:80 ``normalize_input: True`` -- as running_mean / running_var / count)."""
    flat = flat.detach().float().cpu()
    sl, n = mlp_slices(state_dim, 1, hidden)
    assert flat.numel() == n
    sd = {}
    nl = len(hidden)
    for name, off, shp in sl:
        l, kind = int(name[1:]), ("weight" if name[0] == "W" else "bias")
        sd[(f"model.a2c_network.actor_mlp.{2 * l}." if l < nl else "model.a2c_network.value.") + kind] = flat[off:off + _numel(shp)].view(shp).clone()
    if rms is not None:
        mean, var, count = rms
        sd["model.running_mean_std.running_mean"] = mean.detach().double().cpu().clone()
        sd["model.running_mean_std.running_var"] = var.detach().double().cpu().clone()
        sd["model.running_mean_std.count"] = count.detach().double().cpu().reshape(()).clone()
    return sd


def central_value_flat(cv_sd, state_dim, hidden=HIDDEN):
    sl, n = mlp_slices(state_dim, 1, hidden)
    flat = torch.zeros(n)
    nl = len(hidden)
    for name, off, shp in sl:
        l, kind = int(name[1:]), ("weight" if name[0] == "W" else "bias")
        t = _find(cv_sd, (f"actor_mlp.{2 * l}." if l < nl else "value.") + kind)
        if tuple(t.shape) != tuple(shp):
            raise ValueError(f"central value {name}: checkpoint shape {tuple(t.shape)} != network shape {tuple(shp)}")
        flat[off:off + _numel(shp)] = t.detach().float().cpu().reshape(-1)
    rms = None
    if any(k.endswith("running_mean") for k in cv_sd):
        rms = (_find(cv_sd, "running_mean").float(), _find(cv_sd, "running_var").float(), _find(cv_sd, "count").double().reshape(1))
    return flat, rms


def adam_state_dict(m, v, step, slices, lr, order, shapes=None):
    """flat Adam moments -> ``torch.optim.Adam.state_dict()`` with one entry per parameter tensor, in ``order``
    (the order ``model.parameters()`` yields them, which is what the integer ids of a torch optimizer state mean).
    Names that are not in ``slices`` (the a2c net's own, never-trained critic) get zero moments of ``shapes[name]``."""
    by = {n: (o, s) for n, o, s in slices}
    state = {}
    for i, name in enumerate(order):
        if name in by:
            o, s = by[name]
            k = _numel(s)
            ea, es = m[o:o + k].detach().float().cpu().view(s).clone(), v[o:o + k].detach().float().cpu().view(s).clone()
        else:
            ea, es = torch.zeros(shapes[name]), torch.zeros(shapes[name])
        state[i] = {"step": torch.tensor(float(step)), "exp_avg": ea, "exp_avg_sq": es}
    return {"state": state, "param_groups": [{"lr": lr, "betas": (0.9, 0.999), "eps": 1e-08, "weight_decay": 0, "amsgrad": False,
                                             "params": list(range(len(order)))}]}


def a2c_param_entries(in_dim, out_dim, hidden=HIDDEN):
    """[(name, shape)] in the order ``A2CBuilder.Network.parameters()`` yields them -- the integer ids of rl_games' Adam state.
    ``nn.Module.parameters()`` lists a module's OWN parameters first (``sigma``, a root-level nn.Parameter), then its sub-modules in
    registration order: actor_mlp, critic_mlp (``separate: True``), value, mu.  rl_games' optimiser owns all 17 tensors, the
    never-used critic trunk included, and ``A2CBase.set_full_state_weights`` always calls ``optimizer.load_state_dict``: a state
    with any other count or order raises there.  Names: W/b = the actor trunk + mu head of ``ppo.MLP.slices``; ``critic_*`` /
    ``value_*`` = tensors this engine does not train."""
    dims = [in_dim, *hidden]
    n = len(hidden)
    e = [("sigma", (out_dim,))]
    for l in range(n):
        e += [(f"W{l}", (dims[l + 1], dims[l])), (f"b{l}", (dims[l + 1],))]
    for l in range(n):
        e += [(f"critic_W{l}", (dims[l + 1], dims[l])), (f"critic_b{l}", (dims[l + 1],))]
    e += [("value_W", (1, hidden[-1])), ("value_b", (1,))]
    e += [(f"W{n}", (out_dim, hidden[-1])), (f"b{n}", (out_dim,))]
    return e


def actor_param_order(in_dim=None, out_dim=None, hidden=HIDDEN):
    """names of ``a2c_param_entries`` (all 17 tensors); the shapes are only needed to write zero moments for the critic"""
    return [n for n, _ in a2c_param_entries(in_dim or 1, out_dim or 1, hidden)]


LEGACY_ACTOR_ORDER = ["W0", "b0", "W1", "b1", "W2", "b2", "W3", "b3", "sigma"]   # files written by this package before the fix


def save_checkpoint(filename, state):
    """rl_games ``torch_ext.save_checkpoint``: appends '.pth' to the name it is given"""
    if not filename.endswith(".pth"):
        filename = filename + ".pth"
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    torch.save(state, filename)
    return filename


def load_checkpoint(filename):
    """rl_games ``torch_ext.load_checkpoint`` (plain torch.load onto the host)"""
    return torch.load(filename, map_location="cpu", weights_only=False)
