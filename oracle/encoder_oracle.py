"""TEST INFRASTRUCTURE — fp32 CPU oracle of the recompute stage.

The reference computes passage embeddings with sentence-transformers' ``model.encode``
(packages/leann-core/src/leann/embedding_compute.py:231-239); its own explicit equivalent is
HF ``AutoModel`` -> ``last_hidden_state`` -> masked mean pool (:319-335), and the model's
``Normalize`` module L2-normalises (all-MiniLM-L6-v2: mean pool; bge-base-en-v1.5: CLS pool).
``sentence_transformers`` is absent from this image, so the oracle drives the same third-party
arithmetic directly: ``transformers.BertModel`` (transformers 5.5 here; the reference pins
``transformers>=4.30``), eager attention, fp32, batch padded to the longest sequence with an
attention mask exactly as the tokenizer's ``padding=True`` does.

Parity of the encoder is UNPINNED upstream (no golden vectors, no weights in /root/reference):
it is pinned here by this BertModel run on the seeded synthetic weights of leann_b200.synth.
Distances then follow hnsw_embedding_server.py:195-200: ``-E @ q`` or ``sum((E - q)^2)``.
"""
from __future__ import annotations

import numpy as np
import torch


class EncoderOracle:
    def __init__(self, preset, weights: dict[str, np.ndarray], threads: int | None = None):
        from transformers import BertConfig, BertModel

        if threads:
            torch.set_num_threads(threads)
        self.p = preset
        cfg = BertConfig(vocab_size=preset.vocab_size, hidden_size=preset.hidden, num_hidden_layers=preset.layers,
                         num_attention_heads=preset.heads, intermediate_size=preset.ffn, hidden_act="gelu",
                         max_position_embeddings=preset.max_pos, type_vocab_size=preset.type_vocab,
                         layer_norm_eps=preset.ln_eps, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        try:
            cfg._attn_implementation = "eager"  # embedding_compute.py:161
        except Exception:
            pass
        m = BertModel(cfg, add_pooling_layer=False).eval()
        H = preset.hidden
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))
        sd = {
            "embeddings.word_embeddings.weight": t(weights["word_emb"]),
            "embeddings.position_embeddings.weight": t(weights["pos_emb"]),
            "embeddings.token_type_embeddings.weight": t(weights["type_emb"]),
            "embeddings.LayerNorm.weight": t(weights["emb_ln_g"]),
            "embeddings.LayerNorm.bias": t(weights["emb_ln_b"]),
        }
        for l in range(preset.layers):
            wq = weights[f"l{l}.w_qkv"]
            bq = weights[f"l{l}.b_qkv"]
            pre = f"encoder.layer.{l}."
            sd[pre + "attention.self.query.weight"] = t(wq[:H]); sd[pre + "attention.self.query.bias"] = t(bq[:H])
            sd[pre + "attention.self.key.weight"] = t(wq[H:2 * H]); sd[pre + "attention.self.key.bias"] = t(bq[H:2 * H])
            sd[pre + "attention.self.value.weight"] = t(wq[2 * H:]); sd[pre + "attention.self.value.bias"] = t(bq[2 * H:])
            sd[pre + "attention.output.dense.weight"] = t(weights[f"l{l}.w_o"]); sd[pre + "attention.output.dense.bias"] = t(weights[f"l{l}.b_o"])
            sd[pre + "attention.output.LayerNorm.weight"] = t(weights[f"l{l}.ln1_g"]); sd[pre + "attention.output.LayerNorm.bias"] = t(weights[f"l{l}.ln1_b"])
            sd[pre + "intermediate.dense.weight"] = t(weights[f"l{l}.w_1"]); sd[pre + "intermediate.dense.bias"] = t(weights[f"l{l}.b_1"])
            sd[pre + "output.dense.weight"] = t(weights[f"l{l}.w_2"]); sd[pre + "output.dense.bias"] = t(weights[f"l{l}.b_2"])
            sd[pre + "output.LayerNorm.weight"] = t(weights[f"l{l}.ln2_g"]); sd[pre + "output.LayerNorm.bias"] = t(weights[f"l{l}.ln2_b"])
        missing, unexpected = m.load_state_dict(sd, strict=False)
        missing = [k for k in missing if "position_ids" not in k and "token_type_ids" not in k]
        if missing or unexpected:
            raise RuntimeError(f"state dict mismatch: missing={missing} unexpected={unexpected}")
        self.model = m

    @torch.inference_mode()
    def encode_lists(self, seqs: list[np.ndarray], batch_size: int = 32, return_hidden: bool = False) -> np.ndarray:
        out = []
        hid = []
        for b0 in range(0, len(seqs), batch_size):
            chunk = [np.asarray(s[: self.p.max_pos], np.int64) for s in seqs[b0:b0 + batch_size]]
            L = max(len(s) for s in chunk)
            ids = torch.zeros((len(chunk), L), dtype=torch.long)
            mask = torch.zeros((len(chunk), L), dtype=torch.long)
            for i, s in enumerate(chunk):
                ids[i, : len(s)] = torch.from_numpy(s)
                mask[i, : len(s)] = 1
            h = self.model(input_ids=ids, attention_mask=mask).last_hidden_state  # (B, L, H)
            if return_hidden:
                hid += [h[i, : len(s)].numpy().copy() for i, s in enumerate(chunk)]
            if self.p.pooling == 1:
                pooled = h[:, 0]
            else:  # embedding_compute.py:325-332
                mf = mask.unsqueeze(-1).to(h.dtype)
                pooled = (h * mf).sum(1) / mf.sum(1).clamp(min=1)
            if self.p.normalize:
                pooled = torch.nn.functional.normalize(pooled, p=2, dim=1)
            out.append(pooled.numpy().astype(np.float32))
        emb = np.concatenate(out) if out else np.zeros((0, self.p.hidden), np.float32)
        return (emb, hid) if return_hidden else emb

    def encode_store(self, tokens: np.ndarray, offsets: np.ndarray, ids=None, batch_size: int = 32) -> np.ndarray:
        n = len(offsets) - 1
        ids = range(n) if ids is None else ids
        seqs = [tokens[int(offsets[i]): int(offsets[i + 1])] for i in ids]
        return self.encode_lists(seqs, batch_size)

    def distance_fn(self, tokens, offsets, metric_ip: bool = True, batch_size: int = 32):
        """The embedding server's distance branch (hnsw_embedding_server.py:147-211) as a
        callback for the traversal oracle: (q, ids) -> distances."""

        def fn(q, ids):
            E = self.encode_store(tokens, offsets, ids, batch_size)
            if metric_ip:
                return -(E @ q)
            return np.square(E - q.reshape(1, -1)).sum(1)

        return fn
