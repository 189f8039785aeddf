"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference in the build container.

``/root/reference`` needs ``gymnasium`` and ``plotly`` at import time
(diamond/ppo.py:6,10; diamond/utils.py:15-16); neither is installed and neither
is touched by the hot path.  This registers minimal stand-ins in ``sys.modules``
so that ``import diamond`` executes the reference's own, unmodified source.
Used only by ``tests/golden/make_golden.py`` (the GPU box has no /root/reference).
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _install_stubs(obs_dim, act, continuous):
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")
    vector = types.ModuleType("gymnasium.vector")

    class Space:
        pass

    class Box(Space):
        def __init__(self, shape):
            self.shape = tuple(shape)       # only .shape is read (ppo.py:54)

    class Discrete(Space):
        def __init__(self, n):
            self.n = n                      # only .n is read (ppo.py:63)

    class SyncVectorEnv:                    # ctor signature used at ppo.py:124-128
        obs_dim = 4
        act = 2
        continuous = False

        def __init__(self, fns, copy=True, autoreset_mode=None):
            self.single_observation_space = Box((SyncVectorEnv.obs_dim,))
            self.single_action_space = (Box((SyncVectorEnv.act,)) if SyncVectorEnv.continuous
                                        else Discrete(SyncVectorEnv.act))

    SyncVectorEnv.obs_dim, SyncVectorEnv.act, SyncVectorEnv.continuous = obs_dim, act, continuous
    spaces.Space, spaces.Box, spaces.Discrete = Space, Box, Discrete
    gym.Env = type("Env", (), {})
    gym.spaces, gym.vector = spaces, vector
    vector.SyncVectorEnv = SyncVectorEnv

    plotly = types.ModuleType("plotly")
    go = types.ModuleType("plotly.graph_objects")
    pio = types.ModuleType("plotly.io")
    go.layout = type("L", (), {"Template": staticmethod(lambda **kw: kw)})   # utils.py:219
    go.Figure = go.Scatter = go.Bar = object
    pio.templates = {}
    plotly.graph_objects, plotly.io = go, pio
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.vector": vector,
                        "plotly": plotly, "plotly.graph_objects": go, "plotly.io": pio})
    return vector.SyncVectorEnv


def import_reference(obs_dim=4, act=2, continuous=False):
    """Returns the reference ``diamond`` package; (obs_dim, act, continuous) set the stub env's spaces."""
    if "gymnasium" in sys.modules and hasattr(sys.modules["gymnasium"], "vector"):
        sve = sys.modules["gymnasium"].vector.SyncVectorEnv
        sve.obs_dim, sve.act, sve.continuous = obs_dim, act, continuous
    else:
        _install_stubs(obs_dim, act, continuous)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import diamond  # noqa: the reference package (unmodified source)
    assert diamond.__file__.startswith(REFERENCE_ROOT), diamond.__file__
    return diamond


def patch_recurrent_none_checks(diamond):
    """The 2-line fix SURVEY.md §0.4 documents: ``hx or zeros`` -> ``hx if hx is not None else zeros``
    (recurrent_ppo.py:78-79 raise on multi-element tensors).  Everything else is the reference's code."""
    import torch
    from diamond import recurrent_ppo as rp

    def forward(self, x, hx, dones):
        seq_length, batch_size = x.shape[:2]
        if hx is None:
            hx = torch.zeros(1, batch_size, self.hidden_size, dtype=x.dtype, device=x.device)
        if dones is None:
            dones = torch.zeros(seq_length, batch_size, dtype=torch.bool, device=x.device)
        outputs = []
        for t in range(seq_length):
            hx[:, dones[t]] = 0.0
            out, hx = torch.nn.GRU.forward(self, x[t:t + 1], hx)
            outputs.append(out)
        return torch.concatenate(outputs, dim=0), hx

    rp.GRUCore.forward = forward
