"""Drive the UNMODIFIED reference (victorca25/traiNNer) through its own public API.

Used by bench.py's reference arms and by the parity tests.  This module imports NOTHING from
trainner_b200 (the reference arm of the bench must not load the product library) and nothing from
oracle/.  The reference tree is looked up at baseline/_ref/codes (staged by tools/stage_reference.py;
it travels to the GPU box) and, in the build container only, at /root/reference/codes.

Harness-side shims (reference files untouched):
  * matplotlib stub (models/losses.py -> dataops/debug.py -> dataops/flow_utils.py import pyplot);
  * a seeded synthetic torchvision-VGG19 checkpoint pre-seeded under $TORCH_HOME, because
    perceptual.py:141 calls vgg19(pretrained=True) and there is no network;
  * network_G.gaussian False (block.GaussianNoise hard-codes 'cuda' and injects RNG noise,
    block.py:587-598);
  * opt is a NoneDict built by hand (options.parse needs dataset paths).
Reference call path: models/__init__.py:create_model -> models/sr_model.py:SRModel (:17),
feed_data (:115), optimize_parameters (:195); AMP = fp16 autocast + GradScaler
(models/base_model.py:736-744).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.path.join(HERE, "_ref", "codes"), "/root/reference/codes")


def ref_root():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "models")):
            return c
    return None


def reference_available():
    return ref_root() is not None


def install_shims(torch_home):
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not found: run tools/stage_reference.py in the build container "
                           "(baseline/_ref/codes) -- /root/reference does not exist on the GPU box")
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    os.environ["TORCH_HOME"] = torch_home
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def seeded_state(shapes, seed, scale=None):
    """Deterministic weights independent of any module's init order.

    shapes: ordered {key: shape}; each tensor is randn * s with s = scale or sqrt(2/fan_in).
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
            out[k] = torch.randn(shp, generator=g) * 0.1
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            s = scale if scale is not None else (2.0 / fan_in) ** 0.5
            out[k] = torch.randn(shp, generator=g) * s
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
    tmp = path + ".tmp%d" % os.getpid()
    torch.save(sd, tmp)
    os.replace(tmp, path)
    return path


def build_opt(nb=1, hr_size=128, scale=4, use_gan=False, use_fea=False, pixel_weight=1.0,
              feature_weight=1.0, gan_weight=5e-3, upsample_mode="upconv", lr=1e-4, gpu=False,
              use_amp=False, batch_size=1, virtual_batch_size=None, grad_clip=None, grad_clip_value=0.1,
              init_scale=None):
    from options.options import dict_to_nonedict
    from options.defaults import get_network_defaults

    net_g = {"type": "esrgan", "nb": nb, "nf": 64, "gc": 32, "gaussian": False,
             "upsample_mode": upsample_mode}
    opt = {
        "name": "parity",
        "model": "sr",
        "scale": scale,
        "gpu_ids": [0] if gpu else None,
        "is_train": True,
        "use_amp": bool(use_amp),
        "use_swa": False,
        "use_cem": False,
        "datasets": {"train": {"crop_size": hr_size, "batch_size": batch_size,
                               "virtual_batch_size": virtual_batch_size or batch_size, "znorm": False}},
        "path": {"root": "/tmp", "pretrain_model_G": None, "pretrain_model_D": None,
                 "models": "/tmp/_ref_models", "training_state": "/tmp/_ref_state"},
        "network_G": net_g,
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
            "grad_clip": grad_clip, "grad_clip_value": grad_clip_value,
        },
        "logger": {"print_freq": 1},
    }
    if use_gan:
        opt["network_D"] = {"type": "discriminator_vgg"}
    opt = dict_to_nonedict(opt)
    opt = get_network_defaults(opt, True)
    if init_scale is not None:
        # get_network_defaults rebuilds network_G from its known keys (defaults.py:36-63) and drops init_scale;
        # networks.get_network pops it from the final dict (networks.py:116-118), so it is set here
        opt["network_G"]["init_scale"] = init_scale
    return dict_to_nonedict(opt)


def create_reference_model(torch_home="/tmp/_ref_torch_home", seed=0, precision="fp32", **kw):
    """The reference's own SRModel (models/sr_model.py:17) for the given recipe.

    precision: 'fp32' (stock), 'amp' (the reference's native fp16 autocast + GradScaler,
    base_model.py:736-744, via opt.use_amp) or 'bf16' (model.cast = bf16 autocast, amp stays False --
    the like-for-like row SURVEY.md 8d asks for; the two attributes are set from the harness, no
    reference file is touched)."""
    install_shims(torch_home)
    make_vgg19_checkpoint(torch_home)
    import torch

    torch.manual_seed(seed)
    from models import create_model

    opt = build_opt(use_amp=(precision == "amp"), **kw)
    model = create_model(opt)
    if precision == "bf16":
        import functools
        model.cast = functools.partial(torch.autocast, "cuda" if kw.get("gpu") else "cpu", dtype=torch.bfloat16)
        model.amp = False
    elif precision not in ("fp32", "amp"):
        raise ValueError(precision)
    return model, opt


def unwrap(net):
    return net.module if hasattr(net, "module") else net


def perceptual_network(model):
    """the reference's FeatureExtractor instance inside GeneratorLoss (losses.py:654-690)"""
    for l in model.generatorlosses.loss_list:
        if "fea" in l["name"]:
            return l["function"].network
    return None


def time_reference(model, batch, steps, warmup, cuda):
    """ms per feed_data + optimize_parameters (the reference's training iteration, train.py:237-238)."""
    import time

    import torch

    it = 0
    for _ in range(warmup):
        it += 1
        model.feed_data(batch)
        model.optimize_parameters(it)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        it += 1
        model.feed_data(batch)
        model.optimize_parameters(it)
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    return (time.perf_counter() - t0) * 1e3 / steps
