"""Device-side batch packer -- host half.

Stands where the reference's ``SNLIGPT_gen_collate`` (Data/VCRChunkAlign.py:690-741) and the tensor building of the
dataset's ``__getitem__`` (:655-686) stand: takes the ragged per-(question, choice) samples, writes ONE pinned host
buffer (int32 descriptors + one feature block per question), copies it to the device once and lets
``modcr_pack_batch`` (csrc/pack.cu) expand it into the padded batch of the reference's batch contract.  The chunk mask
travels as chunk ids (derived from ``offsets``, :667-670) and is expanded on the device straight into the bit mask the
attention kernels read; the image features of the four choices of a question cross PCIe once, not four times.

``unbatch`` turns a padded synthetic batch (``synth.make_batch``) back into ragged samples -- what a dataset would
hand to the collate function -- for tests and for the end-to-end measurement.
"""
import numpy as np
import torch

from . import ops

I32, I64, F32 = torch.int32, torch.int64, torch.float32


def unbatch(batch, num_choices=4):
    """Padded batch -> list of per-(question, choice) samples with the reference's field names (:683-686).
    The four choices of a question share ONE ``img_feat`` tensor object, as in the reference's ``__getitem__``."""
    N, T = batch["input_ids"].shape
    R = batch["img_feat"].shape[1]
    samples = []
    for n in range(N):
        L = int(batch["input_mask"][n, :T].sum().item())
        q0 = (n // num_choices) * num_choices
        if n == q0:
            feat, fmask = batch["img_feat"][n].clone(), batch["input_mask"][n, T:T + R].clone()
        samples.append(dict(
            input_ids=batch["input_ids"][n, :L].clone(), segment_ids=batch["token_type_ids"][n, :L].clone(),
            input_mask=torch.ones(L), img_feat=feat, img_mask=fmask, target=batch["label"][n].clone(),
            chunk_mask=batch["chunk_attention_mask"][n, :L, :L].clone(), gather_index=batch["gather_index"][n],
            offsets=batch["offsets"][n], total_label=batch["total_label"][n, :L].clone(),
            align_pos=batch["align_pos"][n, :L].clone()))
    return samples


# order of the per-choice tuple the reference's ``__getitem__`` returns (Data/VCRChunkAlign.py:683-686) and the collate
# unzips (:692-693)
REF_FIELDS = ("img_id", "image", "text", "r_input_ids", "r_segment_ids", "r_input_mask", "input_ids", "segment_ids",
              "input_mask", "img_feat", "img_mask", "target", "chunk_mask", "gather_index", "offsets", "ques", "ans",
              "total_label", "align_pos")


def samples_from_reference(inputs):
    """``inputs`` as a DataLoader hands them to ``SNLIGPT_gen_collate``: a list (questions of the batch) of tuples (the
    choices of a question) of 19-tuples -> flat list of sample dicts (``concat`` + ``unzip`` of :692-693)."""
    out = []
    for question in inputs:
        for choice in question:
            if len(choice) != len(REF_FIELDS):
                raise ValueError("expected the %d-field tuple of Data/VCRChunkAlign.py:683-686, got %d fields"
                                 % (len(REF_FIELDS), len(choice)))
            out.append(dict(zip(REF_FIELDS, choice)))
    return out


def pad_language_side(samples):
    """The RoBERTa-side fields of the batch (:695-697): ``pad_sequence(batch_first=True, padding_value=0)`` of
    ``r_input_ids`` / ``r_input_mask`` / ``r_segment_ids``.  Host tensors; ``None`` when the samples carry none."""
    if not samples or samples[0].get("r_input_ids") is None:
        return None
    from torch.nn.utils.rnn import pad_sequence
    return {"r_input_ids": pad_sequence([s["r_input_ids"] for s in samples], batch_first=True, padding_value=0),
            "r_attention_mask": pad_sequence([s["r_input_mask"] for s in samples], batch_first=True, padding_value=0),
            "r_token_type_ids": pad_sequence([s["r_segment_ids"] for s in samples], batch_first=True, padding_value=0)}


class PackedBatch(dict):
    """The batch dict of the reference contract (device tensors) plus ``chunk_bits`` / ``chunk_id`` / ``max_chunks``.
    ``chunk_attention_mask`` holds the int32 bit mask ``[N,T,ceil(T/32)]`` (``SeqBertImgModel`` accepts it in place of
    the dense fp32 ``[N,T,T]`` tensor) and ``gather_index`` the int32 chunk ids ``[N,T]``."""


class BatchPacker:
    def __init__(self, device, feat_dim=2054, pad_T=None, pad_R=None, dense_chunk_mask=False, n_buffers=2):
        """``pad_T`` / ``pad_R``: pad to fixed sizes (static shapes for CUDA-graph replay) instead of the batch maxima the
        reference pads to (:699, :709-711).  ``dense_chunk_mask``: also emit the reference's fp32 ``[N,T,T]`` tensor.
        ``n_buffers``: pinned staging buffers in rotation -- the host may fill buffer i+1 while the asynchronous copy of
        buffer i is still queued; a buffer is rewritten only after the event recorded behind its copy has completed."""
        self.device = torch.device(device)     # host_pack() works anywhere; pack() needs CUDA (no CPU path)
        self.D, self.pad_T, self.pad_R, self.dense = feat_dim, pad_T, pad_R, dense_chunk_mask
        self._slots = [[None, None, None] for _ in range(max(1, int(n_buffers)))]   # [pinned, device buffer, copy event]
        self._next = 0
        self.h2d_bytes = 0

    def _buffer(self, nbytes):
        """Next staging slot ``[pinned, device, event]`` of the rotation, safe to overwrite: waits (host side) for the copy
        that last read the pinned half and for the kernels that last read the device half."""
        slot = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        if slot[2] is not None:
            slot[2].synchronize()
        if slot[0] is None or slot[0].numel() < nbytes:
            slot[0] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            slot[1] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return slot

    def host_pack(self, samples, buf=None):
        """Host half: fills the descriptor buffer (a uint8 tensor of at least the returned size; allocated when ``buf`` is
        None) and returns ``(buf, layout)`` with ``layout = dict(N, T, R, U, hdr_bytes, tok_bytes, feat_off, total,
        max_chunks)``.  Pure CPU work (no CUDA needed): tests decode the buffer with ``oracle/collate.py``."""
        N, D = len(samples), self.D
        lens = [int(s["input_ids"].numel()) for s in samples]
        nbox = [int(s["img_mask"].sum().item()) for s in samples]
        T = max(lens) if self.pad_T is None else self.pad_T
        R = max(nbox) if self.pad_R is None else self.pad_R
        if max(lens) > T or max(nbox) > R:
            raise ValueError("sample longer than pad_T / pad_R")
        # one feature block per distinct img_feat tensor (the choices of a question share theirs)
        blocks, block_of = [], {}
        for s in samples:
            key = s["img_feat"].data_ptr()
            if key not in block_of:
                block_of[key] = len(blocks)
                blocks.append(s["img_feat"])
        U = len(blocks)
        hdr_b, tok_b = N * 4 * 4, N * T * 2 * 4
        feat_off = (hdr_b + tok_b + 15) & ~15
        total = feat_off + U * R * D * 4
        if buf is None:
            buf = torch.empty(total, dtype=torch.uint8)
        nb = buf.numpy()
        header = nb[:hdr_b].view(np.int32).reshape(N, 4)
        tokens = nb[hdr_b:hdr_b + tok_b].view(np.int32).reshape(N, T, 2)
        feats = nb[feat_off:total].view(np.float32).reshape(U, R, D)
        tokens[:] = 0
        max_chunks = 1
        for n, s in enumerate(samples):
            L = lens[n]
            header[n] = (L, nbox[n], int(s["target"]), block_of[s["img_feat"].data_ptr()])
            cid = np.zeros(L, dtype=np.int64)
            offs = s["offsets"]
            k = 1
            for idx, chunk in enumerate(offs):          # gather_index of :667-670, shifted by +1 (0 = no chunk)
                c = len(chunk)
                cid[k:k + c] = idx + 1
                k += c
            if len(offs) > 1023 - 1:
                raise ValueError("more than 1022 chunks in one sequence")
            max_chunks = max(max_chunks, len(offs))
            tl = s["total_label"].numpy().astype(np.int64)
            if tl.size and (tl.min() < 0 or tl.max() >= (1 << 19)):
                raise ValueError("total_label outside [0, 2^19)")
            tokens[n, :L, 0] = s["input_ids"].numpy()
            tokens[n, :L, 1] = (cid | (s["segment_ids"].numpy().astype(np.int64) << 10)
                                | ((s["align_pos"].numpy() != 0).astype(np.int64) << 11) | (tl << 12)).astype(np.int32)
        for u, f in enumerate(blocks):
            r = min(R, f.shape[0])
            feats[u, :r] = f[:r].numpy()
            feats[u, r:] = 0.0
        return buf, dict(N=N, T=T, R=R, U=U, hdr_bytes=hdr_b, tok_bytes=tok_b, feat_off=feat_off, total=total,
                         max_chunks=max_chunks)

    def collate(self, inputs):
        """``collate_fn`` drop-in for ``SNLIGPT_gen_collate`` (run_PMR_ModCR.py:118-120): takes what the DataLoader
        hands over (per-question tuples of the reference's per-choice 19-tuples, CPU tensors) and returns the batch."""
        return self.pack(samples_from_reference(inputs))

    DEVICE_FIELDS = ("input_ids", "token_type_ids", "total_label", "align_pos", "label", "label_long", "input_mask",
                     "gather_index", "chunk_attention_mask", "img_feat")

    def alloc_outputs(self, N, T, R):
        """The device tensors ``pack`` fills.  Allocate once and pass as ``out=`` to keep their addresses fixed -- the
        inputs of a captured CUDA graph (``bench.py``'s end-to-end loop packs a new batch into them every step)."""
        dev, D, cw = self.device, self.D, (T + 31) // 32
        out = PackedBatch(
            input_ids=torch.empty((N, T), dtype=I64, device=dev), token_type_ids=torch.empty((N, T), dtype=I64, device=dev),
            total_label=torch.empty((N, T), dtype=I64, device=dev), align_pos=torch.empty((N, T), dtype=I64, device=dev),
            label=torch.empty((N,), dtype=F32, device=dev), label_long=torch.empty((N,), dtype=I64, device=dev),
            input_mask=torch.empty((N, T + R), dtype=F32, device=dev),
            gather_index=torch.empty((N, T), dtype=I32, device=dev),
            chunk_attention_mask=torch.empty((N, T, cw), dtype=I32, device=dev),
            img_feat=torch.empty((N, R, D), dtype=F32, device=dev))
        if self.dense:
            out["dense_chunk_mask"] = torch.empty((N, T, T), dtype=F32, device=dev)
        return out

    def stage(self, samples, copy_stream=None):
        """Host half + the ONE host-to-device copy of the batch (on ``copy_stream`` when given, so it runs under the
        step that is executing).  Returns a handle for ``expand``."""
        if self.device.type != "cuda":
            raise RuntimeError("modcr_b200: BatchPacker.pack needs a CUDA device (no CPU path)")
        D = self.D
        # size the staging buffers for the worst case of this batch, then fill the pinned one in place
        N = len(samples)
        T_ = max(int(s["input_ids"].numel()) for s in samples) if self.pad_T is None else self.pad_T
        R_ = max(int(s["img_feat"].shape[0]) for s in samples) if self.pad_R is None else self.pad_R
        slot = self._buffer(N * 16 + N * T_ * 8 + 16 + N * R_ * D * 4)
        _, lay = self.host_pack(samples, slot[0])
        total = lay["total"]
        stream = copy_stream if copy_stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(stream):
            slot[1][:total].copy_(slot[0][:total], non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(stream)
        slot[2] = copied              # until ``expand`` replaces it with the event behind the kernels
        self.h2d_bytes = total
        return dict(slot=slot, lay=lay, samples=samples, copied=copied)

    def expand(self, handle, out=None):
        """Device half on the current stream: waits for the copy, then ``modcr_pack_batch`` writes the padded batch."""
        slot, lay, samples = handle["slot"], handle["lay"], handle["samples"]
        N, T, R, D = lay["N"], lay["T"], lay["R"], self.D
        hdr_b, tok_b, feat_off, total, max_chunks = (lay["hdr_bytes"], lay["tok_bytes"], lay["feat_off"], lay["total"],
                                                     lay["max_chunks"])
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(handle["copied"])
        dbuf = slot[1]
        d_header = dbuf[:hdr_b].view(I32)
        d_tokens = dbuf[hdr_b:hdr_b + tok_b].view(I32)
        d_feats = dbuf[feat_off:total].view(F32)
        if out is None:
            out = self.alloc_outputs(N, T, R)
        elif tuple(out["input_ids"].shape) != (N, T) or tuple(out["img_feat"].shape) != (N, R, D):
            raise ValueError("out= tensors were allocated for another batch shape")
        dense = out.get("dense_chunk_mask") if self.dense else None
        ops.pack_batch(d_header, d_tokens, d_feats, N, T, R, D, out["input_ids"], out["input_mask"],
                       token_type_ids=out["token_type_ids"], total_label=out["total_label"], align_pos=out["align_pos"],
                       label=out["label_long"], label_f=out["label"], chunk_id=out["gather_index"],
                       chunk_bits=out["chunk_attention_mask"], chunk_mask=dense, img_feat=out["img_feat"])
        done = torch.cuda.Event()
        done.record(cur)
        slot[2] = done                # the slot may be refilled once these kernels have read the device buffer
        return self._finish(out, samples, max_chunks)

    def pack(self, samples, out=None, copy_stream=None):
        """``stage`` + ``expand``.  ``out``: tensors from ``alloc_outputs`` to fill in place."""
        return self.expand(self.stage(samples, copy_stream), out)

    def _finish(self, out, samples, max_chunks):
        dev = self.device
        out["offsets"] = [s["offsets"] for s in samples]
        out["max_chunks"] = max_chunks
        lang = pad_language_side(samples)                      # language-model inputs of Abstract_Specific (:695-697)
        if lang is not None:
            for k, v in lang.items():
                out[k] = v.to(dev, non_blocking=True)
        out["image"], out["text"] = None, None                 # the CLIP inputs are disabled in the reference too (:725)
        for src, dst in (("img_id", "img_id"), ("ques", "ques_str"), ("ans", "ans_str")):   # pass-through (:726-731)
            if src in samples[0]:
                out[dst] = [s[src] for s in samples]
        return out
