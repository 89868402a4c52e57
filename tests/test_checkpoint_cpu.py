"""rl_games checkpoint layout (seqdex_b200/checkpoint.py) checked on the host against torch modules that
carry rl_games' published module names: a state dict we write must load ``strict=True`` into the network the reference
builds (``A2CBuilder.Network`` under ``a2c_network``; nn_controller.py:55-58, RGC:2098-2106), and reading it back must
return the flat vectors bit for bit."""
import os

import torch

from seqdex_b200 import checkpoint as ck


class _A2CNetwork(torch.nn.Module):
    """name-for-name stand-in of rl_games' actor_critic network, ``separate: True``, fixed sigma (yaml:9-29)"""

    def __init__(self, in_dim, actions, units=(1024, 512, 256), central_value=False):
        super().__init__()
        def trunk():
            layers, d = [], in_dim
            for u in units:
                layers += [torch.nn.Linear(d, u), torch.nn.ELU()]
                d = u
            return torch.nn.Sequential(*layers)
        # registration order of rl_games' A2CBuilder.Network: actor_mlp, critic_mlp, value, then mu and the sigma parameter
        self.actor_mlp = trunk()
        if not central_value:
            self.critic_mlp = trunk()
        self.value = torch.nn.Linear(units[-1], 1)
        if not central_value:
            self.mu = torch.nn.Linear(units[-1], actions)
            self.sigma = torch.nn.Parameter(torch.zeros(actions))


class _Model(torch.nn.Module):
    def __init__(self, net):
        super().__init__()
        self.a2c_network = net


def _flat(n, seed):
    return torch.randn(n, generator=torch.Generator().manual_seed(seed))


def test_actor_state_dict_loads_strictly_and_round_trips():
    sl, n = ck.mlp_slices(396, 23, has_sigma=True)
    assert n == 1_062_656 + 256 * 23 + 23 + 23            # actor trunk + mu head + sigma (SURVEY.md 8e parameter counts)
    flat = _flat(n, 0)
    sd = ck.actor_state_dict(flat, 396, 23, seed=5)
    assert sum(v.numel() for v in sd.values()) == 2_131_503   # the whole a2c net of the reference
    model = _Model(_A2CNetwork(396, 23))
    model.load_state_dict(sd, strict=True)
    # same function: the module evaluates the flat vector's network
    x = torch.randn(4, 396)
    h = x
    for l in range(3):
        (_, ow, sw), (_, ob, sb) = sl[2 * l], sl[2 * l + 1]
        h = torch.nn.functional.elu(h @ flat[ow:ow + sw[0] * sw[1]].view(sw).T + flat[ob:ob + sb[0]])
    (_, ow, sw), (_, ob, sb) = sl[6], sl[7]
    mu = h @ flat[ow:ow + sw[0] * sw[1]].view(sw).T + flat[ob:ob + sb[0]]
    assert torch.allclose(model.a2c_network.mu(model.a2c_network.actor_mlp(x)), mu, atol=1e-5)
    back, critic = ck.actor_flat(model.state_dict(), 396, 23)
    assert torch.equal(back, flat)
    assert critic is not None and len(critic) == 8          # critic trunk (3 x W,b) + value head are carried through
    # ... and survive a re-export untouched
    sd2 = ck.actor_state_dict(back, 396, 23, critic=critic)
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k


def test_actor_flat_rejects_wrong_shapes_and_missing_keys():
    import pytest
    flat = _flat(ck.mlp_slices(396, 23, has_sigma=True)[1], 1)
    sd = ck.actor_state_dict(flat, 396, 23)
    with pytest.raises(ValueError):
        ck.actor_flat(sd, 186, 23)                           # Orient-sized network against a GraspSim file
    del sd["a2c_network.mu.bias"]
    with pytest.raises(KeyError):
        ck.actor_flat(sd, 396, 23)


def test_central_value_round_trip_with_running_mean_std():
    sl, n = ck.mlp_slices(564, 1)
    assert n == 1_234_945
    flat = _flat(n, 2)
    rms = (torch.randn(564), torch.rand(564) + 0.5, torch.tensor([12345.0], dtype=torch.float64))
    sd = ck.central_value_state_dict(flat, 564, rms=rms)

    class _CV(torch.nn.Module):           # CentralValueTrain: network under .model, RunningMeanStd buffers beside it
        def __init__(self):
            super().__init__()
            self.model = _Model(_A2CNetwork(564, 0, central_value=True))
            self.model.running_mean_std = torch.nn.Module()
            for name, shape in (("running_mean", 564), ("running_var", 564), ("count", ())):
                self.model.running_mean_std.register_buffer(name, torch.zeros(shape, dtype=torch.float64))
    cv = _CV()
    cv.load_state_dict(sd, strict=True)
    back, rms2 = ck.central_value_flat(cv.state_dict(), 564)
    assert torch.equal(back, flat)
    assert torch.equal(rms2[0], rms[0]) and torch.equal(rms2[1], rms[1]) and float(rms2[2]) == 12345.0
    # older rl_games releases register the network without the a2c_network wrapper: suffix matching accepts both
    old = {k.replace("model.a2c_network.", "model."): v for k, v in sd.items()}
    back_old, _ = ck.central_value_flat(old, 564)
    assert torch.equal(back_old, flat)


def test_adam_state_dict_is_loadable_by_torch_optim(tmp_path):
    """rl_games' Adam owns ALL parameters of the a2c network in ``parameters()`` order -- sigma first (a root-level nn.Parameter),
    then actor_mlp, critic_mlp, value, mu: 17 tensors -- and ``A2CBase.set_full_state_weights`` always loads the optimiser state,
    so the state we write must load into exactly that optimiser and put every moment on the right tensor."""
    sl, n = ck.mlp_slices(64, 3, hidden=(64, 64, 64), has_sigma=True)
    m, v = _flat(n, 3), _flat(n, 4).abs()
    ent = ck.a2c_param_entries(64, 3, hidden=(64, 64, 64))
    osd = ck.adam_state_dict(m, v, 17, sl, 3e-4, [k for k, _ in ent], shapes=dict(ent))
    net = _A2CNetwork(64, 3, units=(64, 64, 64))
    names = [k for k, _ in net.named_parameters()]
    assert names[0] == "sigma" and names[1].startswith("actor_mlp") and names[7].startswith("critic_mlp") and names[13:] == [
        "value.weight", "value.bias", "mu.weight", "mu.bias"] and len(names) == 17
    assert [tuple(osd["state"][i]["exp_avg"].shape) for i in range(17)] == [tuple(p.shape) for p in net.parameters()]
    opt = torch.optim.Adam(net.parameters(), lr=1.0)
    opt.load_state_dict(osd)                                  # raises on any count / shape mismatch
    assert opt.param_groups[0]["lr"] == 3e-4
    P = dict(net.named_parameters())
    by = {k: (o, shp) for k, o, shp in sl}
    assert float(opt.state[P["sigma"]]["step"]) == 17.0
    assert torch.equal(opt.state[P["sigma"]]["exp_avg"], m[n - 3:])
    o, shp = by["W1"]
    assert torch.equal(opt.state[P["actor_mlp.2.weight"]]["exp_avg"], m[o:o + 64 * 64].view(64, 64))
    o, shp = by["b3"]
    assert torch.equal(opt.state[P["mu.bias"]]["exp_avg_sq"], v[o:o + 3])
    assert float(opt.state[P["critic_mlp.0.weight"]]["exp_avg"].abs().max()) == 0.0     # never trained here: zero moments
    # file round trip through the reference's save / load helpers (torch_ext.save_checkpoint appends '.pth')
    fn = ck.save_checkpoint(os.path.join(tmp_path, "nn", "last_allegro_ep_8"), {"model": ck.actor_state_dict(_flat(ck.mlp_slices(396, 23, has_sigma=True)[1], 9), 396, 23),
                                                                                "optimizer": osd, "epoch": 8})
    assert fn.endswith("last_allegro_ep_8.pth") and os.path.exists(fn)
    got = ck.load_checkpoint(fn)
    assert got["epoch"] == 8 and "a2c_network.sigma" in got["model"]
