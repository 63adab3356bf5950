"""TEST INFRASTRUCTURE ONLY — runs the *unmodified* reference (SilentView/EMCID, /root/reference)
on CPU with synthetic inputs so that golden vectors for the hot path can be generated here.

Nothing under emcid_b200/ may import this module.  It only works in the build container, where
/root/reference exists; on the GPU box tests use the committed fixtures under tests/golden/ and
the independent restatement in oracle/emcid_oracle.py.

Recipe: the reference opens globals.yml relative to the CWD (util/globals.py:8)
and imports diffusers / matplotlib only for type hints and isinstance checks on this path
(util/nethook.py:19,83; emcid/layer_stats.py:12; emcid/compute_z.py:12-15), so those two packages
are replaced by stub modules.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("EMCID_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "emcid"))


class _Dummy:
    def __init__(self, *a, **k):
        pass


def _stub_module(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package

    def __getattr__(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return type(attr, (_Dummy,), {})

    m.__getattr__ = __getattr__
    return m


_imported = None


def import_reference():
    """Import the reference's hot-path modules; returns a namespace with them."""
    global _imported
    if _imported is not None:
        return _imported
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for name in ("diffusers", "diffusers.models", "diffusers.models.attention_processor",
                 "diffusers.models.attention", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _stub_module(name)
    cwd = os.getcwd()
    sys.path.insert(0, REFERENCE_ROOT)
    os.chdir(REFERENCE_ROOT)
    try:
        import util.runningstats as runningstats
        import util.nethook as nethook
        import dsets.stat_dataset as stat_dataset
        import emcid.layer_stats as layer_stats
        import emcid.emcid_main as emcid_main
        import emcid.emcid_hparams as emcid_hparams
    finally:
        os.chdir(cwd)
    _imported = SimpleNamespace(runningstats=runningstats, nethook=nethook, stat_dataset=stat_dataset,
                                layer_stats=layer_stats, emcid_main=emcid_main, emcid_hparams=emcid_hparams)
    return _imported


# ---------------------------------------------------------------------------------------------
# synthetic model / data / tokenizer shared by the harness, the oracle tests and bench.py
# ---------------------------------------------------------------------------------------------
BOS, EOS = 49406, 49407


def make_clip_text_model(kind: str = "clip-l", seed: int = 0, **overrides):
    """Random-init HF CLIPTextModel of the shapes in SURVEY.md Appendix B."""
    from transformers import CLIPTextConfig, CLIPTextModel

    presets = {
        "clip-l": dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                       num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu"),
        "bigg": dict(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                     num_attention_heads=20, max_position_embeddings=77, hidden_act="gelu"),
        "tiny": dict(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=2,
                     num_attention_heads=4, max_position_embeddings=77, hidden_act="quick_gelu"),
        "tiny-gelu": dict(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=2,
                          num_attention_heads=4, max_position_embeddings=77, hidden_act="gelu"),
    }
    cfg = dict(presets[kind])
    cfg.update(overrides)
    vocab = cfg["vocab_size"]
    cfg.setdefault("bos_token_id", vocab - 2)
    cfg.setdefault("eos_token_id", vocab - 1)
    torch.manual_seed(seed)
    model = CLIPTextModel(CLIPTextConfig(**cfg)).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    model.config._name_or_path = f"synthetic/{kind}-seed{seed}"
    return model


def make_captions(n: int, vocab: int, seed: int = 0, min_len: int = 8, max_len: int = 77, full: bool = False):
    """Synthetic token-id captions: BOS first, EOS last, ids ~ U{0..vocab-3}."""
    g = torch.Generator().manual_seed(seed)
    bos, eos = vocab - 2, vocab - 1
    lens = torch.full((n,), max_len) if full else torch.randint(min_len, max_len + 1, (n,), generator=g)
    ids = torch.randint(0, vocab - 2, (n, max_len), generator=g)
    out = []
    for i in range(n):
        L = int(lens[i])
        row = ids[i, :L].clone()
        row[0] = bos
        row[L - 1] = eos
        out.append(row)
    return out


class SynthTokenDataset(torch.utils.data.Dataset):
    """Drop-in for dsets/stat_dataset.py::TokenizedDataset items (:99-110)."""

    def __init__(self, captions):
        self.captions = captions

    def __len__(self):
        return len(self.captions)

    def __getitem__(self, i):
        ids = self.captions[i]
        return dict(input_ids=ids.clone(), position_ids=torch.arange(len(ids)),
                    attention_mask=torch.ones(len(ids), dtype=torch.long))


class FakeTokenizer:
    """Whitespace tokenizer with a hashed vocabulary, enough for find_token_range
    (experiments/causal_trace.py:1057-1103) and tokenize_prompts (emcid/compute_z.py:56-74)."""

    model_max_length = 77

    def __init__(self, vocab_size: int):
        self.vocab_size = vocab_size
        self.bos, self.eos = vocab_size - 2, vocab_size - 1
        self._words = {}

    def _id(self, w: str) -> int:
        import zlib

        i = zlib.crc32(w.encode()) % (self.vocab_size - 2)
        self._words.setdefault(i, w)
        return i

    def encode(self, text, truncation=True, max_length=None):
        ids = [self.bos] + [self._id(w) for w in text.split()] + [self.eos]
        max_length = max_length or self.model_max_length
        if len(ids) > max_length:
            ids = ids[: max_length - 1] + [self.eos]
        return ids

    def __call__(self, prompts, return_tensors="pt", padding=True, truncation=True, max_length=None):
        if isinstance(prompts, str):
            prompts = [prompts]
        enc = [self.encode(p, max_length=max_length if truncation else None) for p in prompts]
        width = max(len(e) for e in enc) if padding is True else (max_length or self.model_max_length)
        ids = torch.full((len(enc), width), self.eos, dtype=torch.long)
        mask = torch.zeros((len(enc), width), dtype=torch.long)
        for i, e in enumerate(enc):
            ids[i, : len(e)] = torch.tensor(e)
            mask[i, : len(e)] = 1
        return {"input_ids": ids, "attention_mask": mask}

    def decode(self, ids):
        if torch.is_tensor(ids):
            ids = ids.tolist()
        if isinstance(ids, int):
            ids = [ids]
        toks = []
        for i in ids:
            i = int(i)  # the reference passes lists of 0-d tensors (experiments/causal_trace.py:1046-1049)
            if i == self.bos:
                toks.append("<|startoftext|>")
            elif i == self.eos:
                toks.append("<|endoftext|>")
            else:
                toks.append(self._words.get(i, f"w{i}"))
        return " ".join(toks)


def make_requests(n: int, seed: int = 0):
    """ICEB-style edit requests (dsets/iceb_dataset.py:325-329 templates)."""
    return [
        {"source": f"artist{i} name{i}", "dest": "art", "seed_train": seed,
         "prompts": ["An image of {}", "A photo of {}", "{}"]}
        for i in range(n)
    ]


@contextlib.contextmanager
def cpu_cuda_patches():
    """emcid_main.py:1072-1073 calls torch.cuda.device / empty_cache unconditionally."""
    dev, ec = torch.cuda.device, torch.cuda.empty_cache
    torch.cuda.device = lambda *_a, **_k: contextlib.nullcontext()
    torch.cuda.empty_cache = lambda: None
    try:
        yield
    finally:
        torch.cuda.device, torch.cuda.empty_cache = dev, ec


def run_reference_layer_stats(model, captions, layer: int, stats_dir: str, sample_size: int,
                              batch_tokens: int = 3072):
    """Reference emcid/layer_stats.py::layer_stats_text_encoder (:140-220), unmodified."""
    ref = import_reference()
    ref.layer_stats.get_ccs_filtered_ds = lambda tokenizer: SynthTokenDataset(captions)
    layer_name = f"text_model.encoder.layers.{layer}.mlp.fc2"
    stat = ref.layer_stats.layer_stats_text_encoder(
        model, None, layer_name, stats_dir=stats_dir, sample_size=sample_size, precision="float32",
        batch_tokens=batch_tokens, progress=lambda x, total=None: x)
    return stat


def write_vstar_cache(cache_name: str, requests, h: int, seed: int = 2):
    g = torch.Generator().manual_seed(seed)
    os.makedirs(os.path.dirname(cache_name) or ".", exist_ok=True)
    vs = []
    for r in requests:
        v = torch.randn(h, generator=g)
        np.savez(cache_name + f"source_{r['source']}_dest_{r['dest']}.npz", v_star=v.numpy())
        vs.append(v)
    return torch.stack(vs, dim=1)  # [h, n]


def make_hparams(layers, mom2_n_samples: int, mom2_update_weight: float = 4000, edit_weight: float = 0.5):
    """Minimal stand-in for emcid/emcid_hparams.py::EMCIDHyperParams (:55-163) with exactly the
    fields the stage-2 loop reads (emcid_main.py:846-1065)."""
    return SimpleNamespace(
        layers=list(layers), mom2_update_weight=mom2_update_weight, edit_weight=edit_weight,
        rewrite_module_tmp="text_model.encoder.layers.{}.mlp.fc2", mom2_dataset="ccs_filtered",
        mom2_n_samples=mom2_n_samples, mom2_dtype="float32", num_edit_tokens=1, objective="ori",
        sld_supervision=False, use_new_compute_z=False, txt_img_align_scale_factor=0)


def run_reference_execute(model, tokenizer, requests, hparams, cache_name: str, stat_dir: str,
                          capture: list | None = None):
    """Reference emcid/emcid_main.py::execute_emcid_text_encoder (:818-1082), unmodified, on CPU.
    When `capture` is a list, the (M, Ks) operands of every torch.linalg.solve call are appended."""
    ref = import_reference()
    ref.emcid_main.COV_CACHE.clear()
    pipe = SimpleNamespace(text_encoder=model, tokenizer=tokenizer, device=next(model.parameters()).device)
    orig_solve = torch.linalg.solve

    def spy(A, B, *a, **k):
        if capture is not None:
            capture.append((A.detach().clone(), B.detach().clone()))
        return orig_solve(A, B, *a, **k)

    torch.linalg.solve = spy
    try:
        with cpu_cuda_patches():
            deltas = ref.emcid_main.execute_emcid_text_encoder(
                pipe, requests, hparams, cache_name=cache_name, verbose=False, stat_dir=stat_dir)
    finally:
        torch.linalg.solve = orig_solve
    return deltas


def run_reference_apply(model, tokenizer, requests, hparams, cache_name: str, stat_dir: str):
    """Reference apply_emcid_to_text_encoder (:769-815); mutates `model` in place."""
    ref = import_reference()
    ref.emcid_main.COV_CACHE.clear()
    pipe = SimpleNamespace(text_encoder=model, tokenizer=tokenizer, device=next(model.parameters()).device)
    with cpu_cuda_patches():
        ref.emcid_main.apply_emcid_to_text_encoder(
            pipe, requests, hparams, device=pipe.device, cache_name=cache_name, stats_dir=stat_dir, verbose=False)
    return model


# ---------------------------------------------------------------------------------------------
# UNet cross-attention K/V modules and the whole-CLIPModel variant (tiny stand-ins)
# ---------------------------------------------------------------------------------------------
class _Attn2(torch.nn.Module):
    def __init__(self, hidden, width):
        super().__init__()
        self.to_k = torch.nn.Linear(hidden, width, bias=False)
        self.to_v = torch.nn.Linear(hidden, width, bias=False)


class _TBlock(torch.nn.Module):
    def __init__(self, hidden, width):
        super().__init__()
        self.attn2 = _Attn2(hidden, width)


class _AttnHolder(torch.nn.Module):
    def __init__(self, hidden, width):
        super().__init__()
        self.transformer_blocks = torch.nn.ModuleList([_TBlock(hidden, width)])


class _Block(torch.nn.Module):
    def __init__(self, hidden, widths):
        super().__init__()
        if widths:
            self.attentions = torch.nn.ModuleList([_AttnHolder(hidden, w) for w in widths])


class TinyUNet(torch.nn.Module):
    """The part of diffusers' UNet2DConditionModel the cross-attention path touches: modules named
    {down,up}_blocks.i.attentions.j.transformer_blocks.0.attn2.to_{k,v} and mid_block.attentions.0... (util/globals.py:37-38)
    that all read `encoder_hidden_states`, plus the config fields the reference's dummy forward needs
    (emcid/layer_stats.py:397-403).  Block layout follows SD-v1.4 in miniature: some blocks have no attention."""

    def __init__(self, hidden=64, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        self.down_blocks = torch.nn.ModuleList([_Block(hidden, [32, 32]), _Block(hidden, [])])
        self.up_blocks = torch.nn.ModuleList([_Block(hidden, []), _Block(hidden, [48])])
        self.mid_block = _Block(hidden, [64])
        self.config = SimpleNamespace(in_channels=4, sample_size=8, _name_or_path="synthetic/tiny-unet")
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, latents, timesteps, encoder_hidden_states=None):
        out = []
        for blocks in (self.down_blocks, self.up_blocks, [self.mid_block]):
            for b in blocks:
                for a in getattr(b, "attentions", []):
                    attn = a.transformer_blocks[0].attn2
                    out.append((attn.to_k(encoder_hidden_states), attn.to_v(encoder_hidden_states)))
        return out


def make_cross_attn_pipe(seed: int = 0, device="cpu"):
    """Stand-in for the StableDiffusionPipeline fields the cross-attention path reads."""
    text = make_clip_text_model("tiny", seed=seed)
    # a freshly initialised final_layer_norm (weight 1, bias 0) makes every output row sum to zero: the second moment of
    # last_hidden_state is then exactly singular along the all-ones direction, and so is lambda*C + K K^T.  Trained encoders
    # have a learned affine there; give the stand-in one too.  The bias has a non-zero mean on purpose: y = w * z + b with
    # sum(z) = 0 puts every row on the hyperplane sum(y_i / w_i) = sum(b_i / w_i), whose distance from the origin sets the
    # smallest eigenvalue of the second moment (a zero-mean bias gave cond(C) = 1e8: keys perturbed by 1e-7 then move the
    # reference's own adj_k by 2e-4, and the fixture would pin rounding noise).
    g = torch.Generator().manual_seed(seed + 7)
    fln = text.text_model.final_layer_norm
    fln.weight.copy_(1.0 + 0.3 * torch.randn(fln.weight.shape, generator=g))
    fln.bias.copy_(0.5 + 0.2 * torch.randn(fln.bias.shape, generator=g))
    text = text.to(device)
    unet = TinyUNet(text.config.hidden_size, seed=seed + 100).to(device)
    return SimpleNamespace(text_encoder=text, unet=unet, tokenizer=FakeTokenizer(text.config.vocab_size),
                           scheduler=SimpleNamespace(config=SimpleNamespace(num_train_timesteps=1000)),
                           device=torch.device(device))


def write_cross_attn_vstar_cache(cache_name: str, requests, pipe, layer_names, seed: int = 2):
    """source_{source}.npz holding one pickled {"v_star": array} per K/V module (emcid/emcid_main.py:365-420)."""
    g = torch.Generator().manual_seed(seed)
    os.makedirs(os.path.dirname(cache_name) or ".", exist_ok=True)
    out = {n: [] for n in layer_names}
    for r in requests:
        payload = {}
        for n in layer_names:
            mod = pipe.unet
            for part in n.split("."):
                mod = getattr(mod, part)
            v = torch.randn(mod.weight.shape[0], generator=g)
            payload[n] = {"v_star": v.numpy()}
            out[n].append(v)
        np.savez(cache_name + f"source_{r['source']}.npz", **payload)
    return {n: torch.stack(v, dim=1) for n, v in out.items()}


def run_reference_cross_attn_stats(pipe, captions, layer_name: str, stats_dir: str, sample_size: int):
    """Reference emcid/layer_stats.py::layer_stats_cross_attn_kv (:333-427), unmodified."""
    ref = import_reference()
    ref.layer_stats.get_ccs_filtered_ds = lambda tokenizer: SynthTokenDataset(captions)
    return ref.layer_stats.layer_stats_cross_attn_kv(pipe, layer_name, stats_dir=stats_dir, sample_size=sample_size,
                                                     precision="float32", progress=lambda x, total=None: x)


def run_reference_cross_attn_edit(pipe, requests, hparams, cache_name: str, stats_dir: str, apply: bool):
    """Reference execute_emcid_cross_attn / apply_emcid_to_cross_attn (emcid/emcid_main.py:314-547); statistics come
    from `stats_dir` (the reference reads its module-level STATS_DIR, :2225)."""
    ref = import_reference()
    ref.emcid_main.COV_CACHE.clear()
    saved = ref.emcid_main.STATS_DIR
    ref.emcid_main.STATS_DIR = stats_dir
    try:
        with cpu_cuda_patches():
            if apply:
                return ref.emcid_main.apply_emcid_to_cross_attn(pipe, requests, hparams, device=pipe.device,
                                                                cache_name=cache_name)
            return ref.emcid_main.execute_emcid_cross_attn(pipe, requests, hparams, cache_name=cache_name, verbose=False)
    finally:
        ref.emcid_main.STATS_DIR = saved


def make_clip_model(seed: int = 0):
    """Tiny transformers.CLIPModel (text tower of the 'tiny' preset + a minimal vision tower) for the
    apply_emcid_to_clip variant (emcid/emcid_main.py:109-311)."""
    from transformers import CLIPConfig, CLIPModel

    text = dict(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                max_position_embeddings=77, hidden_act="quick_gelu", bos_token_id=998, eos_token_id=999)
    vision = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2, image_size=16,
                  patch_size=8)
    torch.manual_seed(seed)
    model = CLIPModel(CLIPConfig(text_config=text, vision_config=vision, projection_dim=32)).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    model.config._name_or_path = f"synthetic/tiny-clipmodel-seed{seed}"
    return model


def text_tower_of(clip_model):
    """A CLIPTextModel sharing the CLIPModel's text weights: the reference's statistics pass calls model(**batch),
    which a whole CLIPModel cannot serve without pixel values, so its statistics files are made from the tower alone."""
    from transformers import CLIPTextModel

    tower = CLIPTextModel(clip_model.config.text_config).eval()
    tower.text_model.load_state_dict(clip_model.text_model.state_dict())
    for p in tower.parameters():
        p.requires_grad_(False)
    tower.config._name_or_path = clip_model.config._name_or_path
    return tower


def run_reference_clip_edit(model, tokenizer, requests, hparams, cache_name: str, stats_dir: str, apply: bool):
    """Reference execute_emcid_clip / apply_emcid_to_clip (emcid/emcid_main.py:109-311) on CPU."""
    ref = import_reference()
    ref.emcid_main.COV_CACHE.clear()
    processor = SimpleNamespace(tokenizer=tokenizer)
    # get_cov_text_encoder's stat_dir default was bound at definition time: point it at the fixture's directory
    fn = ref.emcid_main.get_cov_text_encoder
    saved = fn.__defaults__
    fn.__defaults__ = saved[:-1] + (stats_dir,)
    try:
        with cpu_cuda_patches():
            if apply:
                return ref.emcid_main.apply_emcid_to_clip(model, processor, requests, hparams, device=model.device,
                                                          cache_name=cache_name)
            return ref.emcid_main.execute_emcid_clip(model, processor, requests, hparams, cache_name=cache_name, verbose=False)
    finally:
        fn.__defaults__ = saved
