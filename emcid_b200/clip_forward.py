"""Native text-encoder forward for the statistics pass: thin wrapper over the ``emcid_clip_*`` C ABI.

The reference runs ``model(**batch)`` under ``Trace(..., stop=True)`` for every sub-batch
(emcid/layer_stats.py:210-216); with a HF ``CLIPTextModel`` that forward is the dominant cost of the
pass on a GPU.  ``NativeClipTextEncoder`` copies the weights of such a model into the library once
and then runs embeddings -> [LN1, causal attention, LN2, fc1, act, fc2] on the 3xFP16 tcgen05 GEMM over
packed valid tokens, handing act(fc1) of every edited layer to its ``Mom2Accumulator`` on the device.
Anything that is not a plain fp32 CLIP text tower keeps using the HF forward with the fused kernels
spliced in by hooks (``layer_stats.TextEncoderMom2Pass``) — still no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib


def _text_model(model):
    return getattr(model, "text_model", model)


def text_config(model):
    """The text tower's config: `model.config` for CLIPTextModel(WithProjection), `.text_config` for a whole CLIPModel
    (the reference edits those through apply_emcid_to_clip, emcid/emcid_main.py:109-311)."""
    cfg = model.config
    return getattr(cfg, "text_config", None) or cfg


def supports(model) -> bool:
    """True for an fp32 HF CLIP text tower (CLIPTextModel / CLIPTextModelWithProjection / the text half of CLIPModel)."""
    try:
        tm = _text_model(model)
        emb, layers = tm.embeddings, tm.encoder.layers
        cfg = text_config(model)
        if getattr(cfg, "hidden_act", None) not in ("quick_gelu", "gelu"):
            return False
        if cfg.hidden_size % cfg.num_attention_heads or cfg.hidden_size > 2048 or cfg.max_position_embeddings > 128:
            return False
        if cfg.hidden_size % 4 or cfg.intermediate_size % 4:
            return False
        if cfg.hidden_size // cfg.num_attention_heads not in (16, 32, 64, 128):
            return False
        ly = layers[0]
        for mod in (ly.self_attn.q_proj, ly.self_attn.k_proj, ly.self_attn.v_proj, ly.self_attn.out_proj, ly.mlp.fc1,
                    ly.mlp.fc2):
            if not isinstance(mod, torch.nn.Linear) or mod.weight.dtype != torch.float32:
                return False
        if not isinstance(ly.layer_norm1, torch.nn.LayerNorm) or not isinstance(ly.layer_norm2, torch.nn.LayerNorm):
            return False
        return (isinstance(emb.token_embedding, torch.nn.Embedding) and isinstance(emb.position_embedding, torch.nn.Embedding)
                and emb.token_embedding.weight.is_cuda)
    except AttributeError:
        return False


def pack_batch(batch: Dict[str, torch.Tensor], max_positions: int):
    """Right-padded [B, L] input_ids / position_ids / attention_mask -> packed int32 ids, positions,
    cu_seqlens (on the tensors' device) plus host-side (n_captions, n_tokens).  Returns None when the mask
    is not a right-padding mask (the packed causal forward would then differ from the reference)."""
    ids, mask = batch["input_ids"], batch["attention_mask"]
    pos = batch.get("position_ids")
    if ids.dim() != 2 or ids.numel() == 0:
        return None
    keep = mask != 0
    lengths = keep.sum(dim=1)
    B, L = ids.shape
    right_padded = keep == (torch.arange(L, device=ids.device)[None, :] < lengths[:, None])
    info = torch.stack([right_padded.all().to(torch.int64), lengths.sum(), lengths.max()]).cpu()
    ok, T, longest = bool(info[0]), int(info[1]), int(info[2])
    if not ok or longest > max_positions:
        return None
    cu = torch.zeros(B + 1, dtype=torch.int32, device=ids.device)
    cu[1:] = torch.cumsum(lengths, 0).to(torch.int32)
    if pos is None:
        pos = torch.arange(L, device=ids.device)[None, :].expand(B, L)
    return ids[keep].to(torch.int32), pos[keep].to(torch.int32), cu, B, T


class NativeClipTextEncoder:
    def __init__(self, model, max_tokens: int, max_captions: int):
        if not supports(model):
            raise NotImplementedError("not a plain fp32 HF CLIP text encoder on a CUDA device")
        tm, cfg = _text_model(model), text_config(model)
        self.device = tm.embeddings.token_embedding.weight.device
        self.n_layers = len(tm.encoder.layers)
        self.hidden, self.inter = cfg.hidden_size, cfg.intermediate_size
        self.max_positions = tm.embeddings.position_embedding.weight.shape[0]
        self.max_tokens, self.max_captions = int(max_tokens), int(max_captions)
        lib = _lib.lib()
        self._h = ctypes.c_void_p()
        eps = float(tm.encoder.layers[0].layer_norm1.eps)
        with torch.cuda.device(self.device):
            _lib.check(lib.emcid_clip_create(
                ctypes.byref(self._h), self.device.index, self.n_layers, self.hidden, cfg.num_attention_heads, self.inter,
                _lib.act_code(cfg.hidden_act), self.max_positions, tm.embeddings.token_embedding.weight.shape[0], eps,
                self.max_tokens, self.max_captions))
        self._sig: Dict[object, tuple] = {}   # what was uploaded: (data_ptr, version) of every source tensor
        self._sums: Dict[object, torch.Tensor] = {}   # ... and its content checksums (verify=True catches `.data` writes)
        self.last_sync: Dict[object, list] = {}
        self.keys_token = None                # (caller token, layer) of the last forward_keys call, see compute_ks.py
        self.has_final_norm = False
        self._layer_lists: Dict[int, list] = {}   # the 16 source tensors of every layer, listed once per verified sync
        self._ring = [{"event": None, "bufs": {}} for _ in range(3)]   # pinned staging sets of _to_device
        self._ring_at = 0
        self.sync_weights(model, verify=True)

    @staticmethod
    def _layer_tensors(ly):
        a = ly.self_attn
        mods = [ly.layer_norm1, a.q_proj, a.k_proj, a.v_proj, a.out_proj, ly.layer_norm2, ly.mlp.fc1, ly.mlp.fc2]
        out = []
        for m in mods:
            out.append(m.weight)
            out.append(m.bias)
        return out

    @staticmethod
    def _signature(tensors) -> tuple:
        return tuple((0, 0) if t is None else (t.data_ptr(), t._version) for t in tensors)

    def _checksums(self, tensors) -> torch.Tensor:
        """One int64 per tensor (host): the wrapping sum of its 32-bit words (exact, order independent; any realistic
        edit of a weight changes it), all tensors in ONE launch (emcid_checksum_tensors) and one D2H copy."""
        table = []
        keep = []
        for t in tensors:
            if t is None:
                table += [0, 0]
            else:
                t = t.detach()
                if not t.is_contiguous():
                    t = t.contiguous()
                    keep.append(t)
                table += [t.data_ptr(), t.numel() * t.element_size() // 4]
        table_dev = torch.tensor(table, dtype=torch.int64).to(self.device, non_blocking=True)
        out = torch.empty(len(tensors), dtype=torch.int64, device=self.device)
        _lib.check(_lib.lib().emcid_checksum_tensors(_lib.ptr(table_dev), len(tensors), _lib.ptr(out), self._stream()))
        return out.cpu()

    def invalidate(self) -> None:
        """Forget what was uploaded: the next sync re-uploads everything it needs."""
        self._sig.clear()
        self._sums.clear()
        self._layer_lists.clear()
        self.keys_token = None

    def sync_weights(self, model, upto_layer: Optional[int] = None, verify: bool = False) -> int:
        """(Re-)upload the embeddings, the final layer norm and layers [0, upto_layer] whose source tensors changed since
        the last upload.  Cheap test: (data_ptr, `Tensor._version`) — PyTorch bumps the version on every in-place write,
        which is how the edit loop's `w[...] = w0 + dW` (reference emcid_main.py:1061) becomes visible here.  Writes
        through `param.data` do NOT bump it (and a replaced Parameter can reuse a pointer): `verify=True` additionally
        compares a content checksum of every source tensor (one small reduction per tensor, one D2H of the results; the
        edit loop asks for it once per edit, `invalidate()` forces a full upload).  Returns the number of layers
        uploaded; `self.last_sync` = {layer, "emb" or "final_norm": indices of the source tensors that had changed}."""
        tm = _text_model(model)
        lib = _lib.lib()
        n = 0
        self.last_sync = {}
        last = self.n_layers - 1 if upto_layer is None else min(int(upto_layer), self.n_layers - 1)
        groups = {"emb": [tm.embeddings.token_embedding.weight, tm.embeddings.position_embedding.weight]}
        fln = getattr(tm, "final_layer_norm", None)
        if isinstance(fln, torch.nn.LayerNorm) and fln.weight is not None and fln.bias is not None:
            groups["final_norm"] = [fln.weight, fln.bias]
        # the 16 source tensors of a layer, listed once per verified sync (a replaced Parameter object is picked up there;
        # between two of them — inside one edit — weights only change in place): 12 layers x 16 attribute walks per call
        # were 0.8 ms of every 100-concept edit
        cache = self._layer_lists
        for i in range(last + 1):
            if verify or i not in cache:
                cache[i] = self._layer_tensors(tm.encoder.layers[i])
            groups[i] = cache[i]
        stale: Dict[object, list] = {}       # group -> indices of the tensors whose content changed behind the version counter
        if verify:
            with torch.cuda.device(self.device):
                flat = self._checksums([t for v in groups.values() for t in v])
            at = 0
            for k, v in groups.items():
                cur = flat[at: at + len(v)]
                at += len(v)
                old = self._sums.get(k)
                if old is not None and not torch.equal(old, cur):
                    stale[k] = [j for j in range(len(v)) if int(old[j]) != int(cur[j])]
                self._sums[k] = cur
        with torch.cuda.device(self.device):
            stream = _lib.current_stream_ptr()
            emb = groups["emb"]
            sig = self._signature(emb)
            if "final_norm" in groups:
                fsig = self._signature(groups["final_norm"])
                if self._sig.get("final_norm") != fsig or "final_norm" in stale:
                    self.last_sync["final_norm"] = [0, 1]
                    w, b = (t.detach().contiguous() for t in groups["final_norm"])
                    _lib.check(lib.emcid_clip_set_final_norm(self._h, _lib.ptr(w), _lib.ptr(b), stream))
                    self._sig["final_norm"] = fsig
                    self.has_final_norm = True
                    n += 1
            if self._sig.get("emb") != sig or "emb" in stale:
                self.last_sync["emb"] = [0, 1]
                tok, pos = emb[0].detach().contiguous(), emb[1].detach().contiguous()
                _lib.check(lib.emcid_clip_set_embeddings(self._h, _lib.ptr(tok), _lib.ptr(pos), stream))
                self._sig["emb"] = sig
            for i in range(last + 1):
                src = groups[i]
                sig = self._signature(src)
                if self._sig.get(i) == sig and i not in stale:
                    continue
                old = self._sig.get(i)
                self.last_sync[i] = (list(range(16)) if old is None else
                                     sorted(set(j for j in range(16) if old[j] != sig[j]) | set(stale.get(i, ()))))
                tensors = [None if t is None else t.detach().contiguous() for t in src]
                arr = (ctypes.c_void_p * 16)(*[_lib.ptr(t) or None for t in tensors])
                mask = 0
                for j in self.last_sync[i]:
                    mask |= 1 << j
                _lib.check(lib.emcid_clip_update_layer(self._h, i, arr, mask, stream))   # first upload: all 16 bits
                self._sig[i] = sig
                n += 1
            # no device synchronisation: the library's copies / splits are queued on the current stream, where every later
            # write to (or release of) a source tensor is ordered behind them
        return n

    def _stream(self) -> int:
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    def _to_device(self, *tensors):
        """Packed host tensors of one block -> device, asynchronously.  The blocks come out of `torch.cat` on the host
        (stat_dataset.PackedReblocker) in pageable memory, and a pageable host-to-device copy is synchronous with the
        stream: it only starts once the previous block's kernels have drained, so the device then idles through the copy, the
        first launches and whatever the host does before them (loader wait, Python) — 0.27 s of a 4.9 s pass on a slow host.
        A ring of three pinned staging sets (an event per set marks the copies done) makes the copies asynchronous and lets
        the host queue up to three blocks ahead of the device."""
        if all(t.is_cuda for t in tensors):
            return tuple(t.to(self.device, non_blocking=True) for t in tensors)
        slot = self._ring[self._ring_at]
        self._ring_at = (self._ring_at + 1) % len(self._ring)
        if slot["event"] is not None:
            slot["event"].synchronize()
        out = []
        for i, t in enumerate(tensors):
            if t.is_cuda or t.is_pinned():
                out.append(t.to(self.device, non_blocking=True))
                continue
            flat = t.reshape(-1)
            buf = slot["bufs"].get(i)
            if buf is None or buf.dtype != flat.dtype or buf.numel() < flat.numel():
                buf = slot["bufs"][i] = torch.empty(max(flat.numel(), 1024), dtype=flat.dtype, pin_memory=True)
            view = buf[: flat.numel()]
            view.copy_(flat)
            out.append(view.to(self.device, non_blocking=True).reshape(t.shape))
        slot["event"] = torch.cuda.Event()
        slot["event"].record(torch.cuda.current_stream(self.device))
        return tuple(out)

    def _run(self, ids, pos, cu, S, T, n_layers, stat_layers, accs, hidden_out):
        ids, pos, cu = self._to_device(ids, pos, cu)
        n_stat = len(stat_layers)
        layers_arr = (ctypes.c_int * max(n_stat, 1))(*stat_layers)
        accs_arr = (ctypes.c_void_p * max(n_stat, 1))(*[a._h for a in accs])
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_clip_forward(self._h, _lib.ptr(ids), _lib.ptr(pos), _lib.ptr(cu), S, T, n_layers, n_stat,
                                                     layers_arr, accs_arr, _lib.ptr(hidden_out), self._stream()))
        s = torch.cuda.current_stream(self.device)
        for t in (ids, pos, cu):
            t.record_stream(s)

    def forward_stats(self, ids, pos, cu, S: int, T: int, stat_layers: Sequence[int], accs) -> None:
        """mom2 / count of every layer in `stat_layers` (ascending) += statistics of the packed tokens."""
        order = sorted(range(len(stat_layers)), key=lambda i: stat_layers[i])
        self._run(ids, pos, cu, S, T, 0, [int(stat_layers[i]) for i in order], [accs[i] for i in order], None)

    def forward_hidden(self, ids, pos, cu, S: int, T: int, n_layers: Optional[int] = None) -> torch.Tensor:
        """Residual stream [T, hidden] after `n_layers` full layers (== HF hidden_states[n_layers], packed)."""
        n_layers = self.n_layers if n_layers is None else n_layers
        out = torch.empty(T, self.hidden, dtype=torch.float32, device=self.device)
        self._run(ids, pos, cu, S, T, n_layers, [], [], out)
        return out

    def forward_final(self, ids, pos, cu, S: int, T: int, acc=None, rows: Optional[torch.Tensor] = None,
                      want_all: bool = False) -> Optional[torch.Tensor]:
        """The text encoder's output last_hidden_state = final_layer_norm(all layers) over the packed tokens
        (emcid_clip_forward_final): `acc` (a Mom2Accumulator with d == hidden) += its second moment and token count;
        `rows` (packed token indices) or `want_all` return the fp32 rows [n, hidden]."""
        if not getattr(self, "has_final_norm", False):
            raise NotImplementedError("this text model has no final_layer_norm to read last_hidden_state from")
        ids, pos, cu = self._to_device(ids, pos, cu)
        out = None
        n_rows = 0
        if rows is not None:
            rows = rows.to(self.device, dtype=torch.int32, non_blocking=True).contiguous()
            n_rows = rows.numel()
        elif want_all:
            n_rows = T
        if n_rows:
            out = torch.empty(n_rows, self.hidden, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_clip_forward_final(self._h, _lib.ptr(ids), _lib.ptr(pos), _lib.ptr(cu), S, T,
                                                           acc._h if acc is not None else None, _lib.ptr(rows), n_rows,
                                                           _lib.ptr(out), self._stream()))
        s = torch.cuda.current_stream(self.device)
        for t in (ids, pos, cu) + ((rows,) if rows is not None else ()):
            t.record_stream(s)
        self.keys_token = None
        return out

    def forward_keys(self, ids, pos, cu, S: int, T: int, layer: int, rows: torch.Tensor,
                     resume_layer: int = -1) -> Tuple[torch.Tensor, torch.Tensor]:
        """fc2 input [R, intermediate] and fc2 output [R, hidden] of encoder layer `layer` at the packed token rows
        `rows` (int32), see emcid_clip_forward_keys in include/emcid_b200.h (resume_layer: continue from the previous
        keys call at that layer over the same tokens)."""
        ids = ids.to(self.device, non_blocking=True)
        pos = pos.to(self.device, non_blocking=True)
        cu = cu.to(self.device, non_blocking=True)
        rows = rows.to(self.device, dtype=torch.int32, non_blocking=True).contiguous()
        R = rows.numel()
        k = torch.empty(R, self.inter, dtype=torch.float32, device=self.device)
        z = torch.empty(R, self.hidden, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_clip_forward_keys(self._h, _lib.ptr(ids), _lib.ptr(pos), _lib.ptr(cu), S, T, int(layer),
                                                          _lib.ptr(rows), R, _lib.ptr(k), _lib.ptr(z), int(resume_layer),
                                                          self._stream()))
        s = torch.cuda.current_stream(self.device)
        for t in (ids, pos, cu, rows):
            t.record_stream(s)
        return k, z

    PROFILE_TAGS = ("qkv", "out_proj", "fc1", "fc1_edited", "fc2", "attention", "layernorm")

    def profile(self, enable: bool = True) -> None:
        _lib.check(_lib.lib().emcid_clip_profile(self._h, 1 if enable else 0))

    def get_profile(self) -> Dict[str, Dict[str, float]]:
        """{kernel class: {launches, ms, flops}} since the last call (see emcid_clip_get_profile)."""
        out = (ctypes.c_double * (3 * len(self.PROFILE_TAGS)))()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_clip_get_profile(self._h, out))
        return {t: {"launches": out[3 * i], "ms": out[3 * i + 1], "flops": out[3 * i + 2]}
                for i, t in enumerate(self.PROFILE_TAGS)}

    def launches(self) -> int:
        return int(_lib.lib().emcid_clip_launches(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().emcid_clip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------
# one cached native encoder per live model for the key extraction of the edit (emcid_b200/compute_ks.py)
# ---------------------------------------------------------------------------------------------------------
_KEY_ENCODERS: Dict[int, Tuple["weakref.ref", NativeClipTextEncoder]] = {}


def release_key_encoders() -> None:
    for _, enc in list(_KEY_ENCODERS.values()):
        enc.close()
    _KEY_ENCODERS.clear()


def key_encoder(model, n_tokens: int, n_captions: int, upto_layer: int, verify: bool = False
                ) -> Optional[NativeClipTextEncoder]:
    """The cached encoder of `model`, grown to the requested capacity, with the weights of layers [0, upto_layer]
    brought up to date (`verify`: by content checksum as well, see NativeClipTextEncoder.sync_weights).  None when the
    model is not a plain fp32 CLIP text tower on a CUDA device or EMCID_NATIVE_KEYS=0."""
    if os.environ.get("EMCID_NATIVE_KEYS", "1") == "0" or not supports(model):
        return None
    key = id(model)
    hit = _KEY_ENCODERS.get(key)
    enc = None
    if hit is not None:
        ref, enc = hit
        if ref() is not model:        # id reuse after the old model died
            enc.close()
            enc = None
    if enc is not None and (enc.max_tokens < n_tokens or enc.max_captions < n_captions):
        n_tokens, n_captions = max(n_tokens, enc.max_tokens), max(n_captions, enc.max_captions)
        torch.cuda.synchronize(enc.device)
        enc.close()
        enc = None
    if enc is None:
        enc = NativeClipTextEncoder(model, max(int(n_tokens), 1024), max(int(n_captions), 64))
        try:
            ref = weakref.ref(model, lambda _r, k=key: _drop_key_encoder(k))
        except TypeError:
            ref = (lambda m=model: m)
        _KEY_ENCODERS[key] = (ref, enc)
    else:
        enc.sync_weights(model, upto_layer, verify=verify)
    return enc


def _drop_key_encoder(key: int) -> None:
    hit = _KEY_ENCODERS.pop(key, None)
    if hit is not None:
        try:
            hit[1].close()
        except Exception:
            pass
