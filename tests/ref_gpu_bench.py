"""Yardstick (diagnostic, not the bench line): the reference ALGORITHM as plain PyTorch on the GPU -- the oracle's modules on cuda under
torch.autocast(bfloat16), autograd backward, clip_grad_norm_, and the reference's per-tensor AdamW loop -- i.e. what the reference's own
PyTorch path (cuBLAS / SDPA-less attention / ATen elementwise) achieves on the same B200.  Prints pairs/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import clip_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = dict(O.vit_b16_bert_base_config(), text_attention_probs_dropout_prob=0.0, text_hidden_dropout_prob=0.0)
sd = {k: v.cuda() for k, v in O.init_state_dict(cfg, seed=1234).items()}
pixels, ids = O.synthetic_batch(cfg, B, seq_len=77, seed=1234)
pixels, ids = pixels.cuda(), ids.cuda()
names = O.trainable_names(sd)
params = {k: sd[k].clone().requires_grad_(True) for k in names}
state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}
full = dict(sd); full.update(params)


def step(i):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = O.clip_forward(full, cfg, pixels, ids)
    loss = O.clip_loss(out["logits_per_text"].float())
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    gl = [g for g in grads if g is not None]
    torch.nn.utils.clip_grad_norm_(gl, 1.0) if False else O.clip_grad_norm(gl, 1.0)
    with torch.no_grad():
        for k, g in zip(names, grads):
            if g is None:
                continue
            m, v = state[k]
            O.adamw_step(params[k], g, m, v, i + 1, 1e-5, 1e-4 if O.uses_weight_decay(k) else 0.0)
    return loss


for i in range(3):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
for i in range(n):
    l = step(3 + i)
l.item()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"reference-algorithm PyTorch GPU path (bf16 autocast, B={B}): {dt*1e3:.1f} ms/step  {B/dt:.0f} pairs/s  loss {l.item():.4f}")
