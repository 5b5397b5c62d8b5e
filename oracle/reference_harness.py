"""Import the UNMODIFIED reference modules from /root/reference (test infrastructure; works only where that tree
exists, i.e. in the build container - never on the GPU box, never from the product).

The reference imports several packages that are absent here and unused on the decode / sampling paths; they are
replaced by empty stub modules: omegaconf, pytorch_lightning(+utilities.distributed), pytorch_memlab,
icecream, taming.modules.vqvae.quantize.
"""
import os
import sys
import types

import torch.nn as nn

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "ldm"))


def install_stubs():
    sys.dont_write_bytecode = True   # /root/reference is read-only

    def stub(name, **attrs):
        if name in sys.modules and not getattr(sys.modules[name], "_alcm_stub", False):
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m._alcm_stub = True
        sys.modules[name] = m
        return m

    class _OC:
        @staticmethod
        def create(x=None):
            return x

    stub("omegaconf", OmegaConf=_OC, ListConfig=list, DictConfig=dict)
    stub("pytorch_lightning", LightningModule=nn.Module, Callback=object, Trainer=object, seed_everything=lambda *a, **k: None)
    stub("pytorch_lightning.utilities", distributed=None)
    stub("pytorch_lightning.utilities.distributed", rank_zero_only=lambda f: f)
    stub("pytorch_memlab", profile=lambda f: f, LineProfiler=object)
    stub("icecream", ic=lambda *a, **k: None)
    stub("taming")
    stub("taming.modules")
    stub("taming.modules.vqvae")
    stub("taming.modules.vqvae.quantize", VectorQuantizer2=object, VectorQuantizer=object)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def build_lcm_audio():
    """The real LCM_audio of configs/audiolcm.yaml (random init, no checkpoints, text encoder replaced by the
    reference's own '__is_unconditional__' switch, lcm_audio.py:515-518).  ~260 M parameters, a few seconds on CPU."""
    import yaml
    install_stubs()
    from ldm.models.diffusion.lcm_audio import LCM_audio
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "audiolcm.yaml")))["model"]["params"]
    cfg.pop("ckpt_path", None)
    cfg.pop("scheduler_config", None)
    cfg["first_stage_config"]["params"].pop("ckpt_path", None)
    cfg["first_stage_config"]["params"]["lossconfig"] = {"target": "torch.nn.Identity"}
    cfg["cond_stage_config"] = "__is_unconditional__"
    ddconfig = cfg["first_stage_config"]["params"]["ddconfig"]
    key = cfg["conditioning_key"]
    model = LCM_audio(**cfg).eval()
    # '__is_unconditional__' also clears the wrappers' conditioning_key (ddpm.py); restore the yaml's value ('crossattn')
    # so that apply_model routes the (stubbed) text context to the DiT exactly as in the shipped configuration
    for w in (model.model, model.unet, model.target_unet):
        w.conditioning_key = key
    return model, ddconfig
