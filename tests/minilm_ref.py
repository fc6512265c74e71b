"""PyTorch f32 reference of the MiniLM-L6-v2 lane (test infrastructure): transformers.BertModel with
the all-MiniLM-L6-v2 geometry and SEEDED SYNTHETIC weights (no model files exist offline; the
reference itself pins MiniLM values only through a SHA-256 digest,
crates/frankensearch-embed/src/model_manifest.rs:308-314), followed by the pooling contract of
crates/frankensearch-embed/src/fastembed_embedder.rs:317-353, :416-426."""
import numpy as np
import torch


def make_bert(seed=0, vocab=2000, layers=6, max_pos=512):
    from transformers import BertConfig, BertModel

    torch.manual_seed(seed)
    cfg = BertConfig(vocab_size=vocab, hidden_size=384, num_hidden_layers=layers, num_attention_heads=12,
                     intermediate_size=1536, max_position_embeddings=max_pos, type_vocab_size=2,
                     layer_norm_eps=1e-12, hidden_act="gelu", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    model = BertModel(cfg, add_pooling_layer=False).eval()
    # default init leaves biases at 0 and LayerNorm at identity: perturb so every term is exercised,
    # and widen the linear weights so attention is not uniform
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("LayerNorm.weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif "embeddings" in name:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.06 * torch.randn(p.shape, generator=g))
    return model


def reference_embed(model, batches):
    lens = [len(x) for x in batches]
    t = max(max(lens), 1)
    ids = torch.zeros((len(batches), t), dtype=torch.long)
    mask = torch.zeros((len(batches), t), dtype=torch.long)
    for i, x in enumerate(batches):
        ids[i, :len(x)] = torch.tensor(x, dtype=torch.long)
        mask[i, :len(x)] = 1
    with torch.no_grad():
        h = model(input_ids=ids, attention_mask=mask, token_type_ids=torch.zeros_like(ids)).last_hidden_state
        m = mask.unsqueeze(-1).to(h.dtype)
        pooled = (h * m).sum(1) / m.sum(1).clamp(min=1e-9)
        v = torch.nn.functional.normalize(pooled, p=2, dim=1, eps=1e-12)   # fastembed normalize
        n2 = (v * v).sum(1, keepdim=True)                                     # adapter normalize_in_place
        ok = torch.isfinite(n2) & (n2 > torch.finfo(torch.float32).eps)
        v = torch.where(ok, v / n2.sqrt(), torch.zeros_like(v))
    out = v.numpy().astype(np.float32)
    for i, n in enumerate(lens):
        if n == 0:
            out[i] = 0.0
    return out


def state_dict_numpy(model):
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in model.state_dict().items()}
