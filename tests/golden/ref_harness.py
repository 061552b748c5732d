"""Harness that imports the UNMODIFIED reference (read-only, /root/reference/codes) on CPU.

Only used in the build container to (a) validate oracle/ against the real reference and
(b) generate the committed golden fixtures (make_golden.py).  /root/reference does not exist on
the GPU box, so nothing under tests/ imports this module at test time except behind
`pytest.importorskip`-style guards.

Shims (all harness-side; reference files untouched) -- see SURVEY.md section 8c:
  * matplotlib stub (models/losses.py -> dataops/debug.py -> dataops/flow_utils.py imports pyplot)
  * seeded synthetic VGG19 checkpoint pre-seeded under $TORCH_HOME (perceptual.py:141 downloads)
  * network_G.gaussian False (block.GaussianNoise hard-codes cuda, block.py:592)
  * opt is a NoneDict built by hand (no dataset paths)
"""
import os
import sys
import types

REF_ROOT = "/root/reference/codes"


def reference_available():
    return os.path.isdir(REF_ROOT)


def _install_shims(torch_home):
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    os.environ["TORCH_HOME"] = torch_home
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def seeded_state(shapes, seed, scale=None):
    """Deterministic weights independent of any module's init order.

    shapes: ordered {key: shape}; each tensor is randn * s with s = scale or 1/sqrt(fan_in).
    BatchNorm: weight ~ 1 + 0.1 randn, bias 0.1 randn, running_mean 0, running_var 1.
    """
    import torch

    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        shp = tuple(shp)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            out[k] = torch.ones(shp)
        elif len(shp) == 1:
            v = torch.randn(shp, generator=g) * 0.1
            # BN weight (paired with running stats) is centred at 1; conv/linear bias at 0
            out[k] = v
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            s = scale if scale is not None else (2.0 / fan_in) ** 0.5
            out[k] = torch.randn(shp, generator=g) * s
    # BN weights: keys whose sibling running_mean exists
    for k in list(out.keys()):
        if k.endswith(".weight") and (k[: -len("weight")] + "running_mean") in out:
            out[k] = out[k] + 1.0
    return out


def make_vgg19_checkpoint(torch_home, seed=7):
    """Write a seeded synthetic torchvision-VGG19 checkpoint where perceptual.py:141 looks for it."""
    import torch
    import torchvision

    path = os.path.join(torch_home, "hub", "checkpoints", "vgg19-dcbb9e9d.pth")
    if os.path.exists(path):
        return path
    os.makedirs(os.path.dirname(path), exist_ok=True)
    net = torchvision.models.vgg19(weights=None)
    shapes = {k: v.shape for k, v in net.state_dict().items()}
    sd = seeded_state(shapes, seed)
    torch.save(sd, path)
    return path


def build_opt(nb=1, hr_size=128, scale=4, use_gan=False, use_fea=False, pixel_weight=1.0,
              feature_weight=1.0, gan_weight=5e-3, upsample_mode="upconv", lr=1e-4):
    from options.options import dict_to_nonedict
    from options.defaults import get_network_defaults

    opt = {
        "name": "golden",
        "model": "sr",
        "scale": scale,
        "gpu_ids": None,
        "is_train": True,
        "use_amp": False,
        "use_swa": False,
        "use_cem": False,
        "datasets": {"train": {"crop_size": hr_size, "batch_size": 1, "virtual_batch_size": 1,
                               "znorm": False}},
        "path": {"root": "/tmp", "pretrain_model_G": None, "pretrain_model_D": None,
                 "models": "/tmp/_golden_models", "training_state": "/tmp/_golden_state"},
        "network_G": {"type": "esrgan", "nb": nb, "nf": 64, "gc": 32, "gaussian": False,
                      "upsample_mode": upsample_mode},
        "train": {
            "lr_G": lr, "lr_D": lr, "optim_G": "adam", "optim_D": "adam",
            "beta1_G": 0.9, "beta2_G": 0.999, "beta1_D": 0.9, "beta2_D": 0.999,
            "weight_decay_G": 0, "weight_decay_D": 0,
            "lr_scheme": "MultiStepLR", "lr_steps": [10 ** 9], "lr_gamma": 0.5,
            "pixel_criterion": "l1", "pixel_weight": pixel_weight,
            "feature_criterion": "l1" if use_fea else None,
            "feature_weight": feature_weight if use_fea else 0,
            "gan_type": "vanilla" if use_gan else None,
            "gan_weight": gan_weight if use_gan else 0,
            "D_update_ratio": 1, "D_init_iters": 0,
            "niter": 10 ** 9,
        },
        "logger": {"print_freq": 1},
    }
    if use_gan:
        opt["network_D"] = {"type": "discriminator_vgg"}
    opt = dict_to_nonedict(opt)
    opt = get_network_defaults(opt, True)
    return dict_to_nonedict(opt)


def create_reference_model(torch_home="/tmp/_golden_torch_home", **kw):
    _install_shims(torch_home)
    make_vgg19_checkpoint(torch_home)
    import torch

    torch.manual_seed(0)
    from models import create_model

    opt = build_opt(**kw)
    model = create_model(opt)
    return model, opt
