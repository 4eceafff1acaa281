"""Golden vectors for the fused cross-entropy: nn.CrossEntropyLoss(ignore_index=pad) exactly as train.py:124 builds
it, and the reference's own utils.accuracy (imported unmodified from /root/reference/src/utils.py).
Output: tests/golden/ce_*.npz.  Run in the build container."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from utils import accuracy  # noqa: E402  (the reference's function)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def case(name, M, V, seed, scale, pad_frac):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(M, V, generator=g) * scale).to(torch.bfloat16).float()   # what the bf16 head hands over
    target = torch.randint(1, V, (M,), generator=g)
    target[torch.rand(M, generator=g) < pad_frac] = 0
    # make the target the best / a top-5 logit now and then so that the hit counts are not trivially zero
    for i in range(0, M, 3):
        logits[i, target[i]] = logits[i].max() + (1.0 if i % 2 == 0 else -0.01)
    x = logits.clone().requires_grad_(True)
    ce = torch.nn.CrossEntropyLoss(ignore_index=0)                                   # train.py:124
    loss = ce(x, target)
    loss.backward()
    accs = accuracy(logits, target, topk=(1, 5), ignore_index=0)                     # train.py:256
    n = int((target != 0).sum())
    np.savez_compressed(os.path.join(OUT, f"ce_{name}.npz"), logits=logits.numpy(), target=target.numpy(),
                        loss=np.float32(loss.item()), grad=x.grad.numpy(), count=np.int32(n),
                        top1=np.int32(round(accs[1] * n)), top5=np.int32(round(accs[5] * n)))
    print(name, float(loss), accs, n)


if __name__ == "__main__":
    case("v1007", 96, 1007, 1, 2.0, 0.2)
    case("v1017", 64, 1017, 2, 4.0, 0.0)
    case("v300", 50, 300, 3, 1.0, 0.5)
