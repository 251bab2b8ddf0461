"""Deterministic "random-init" weights for benchmarks and parity tests.

The reference zero-initialises every ResBlock's second conv, every attention ``proj_out``
and the final conv (``dlpm/models/unet.py:156-158,215,435`` via ``nn.py:68-74``), so a freshly
constructed UNet outputs exactly 0 and any parity/throughput check on it would be vacuous
(hard part 1).  ``randomize_parameters_`` re-draws EVERY parameter with O(1)
activation scale.  Each tensor is seeded by (seed, crc32(parameter name)), so any module
with the same parameter names and shapes -- the reference's ``UNetModel``/``MLPModel`` or this
package's mirrors -- gets bit-identical weights regardless of registration order.
"""
import zlib

import torch


def randomize_parameters_(module, seed: int):
    with torch.no_grad():
        for name, p in module.named_parameters():
            g = torch.Generator().manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
            if p.dim() >= 2:
                fan_in = p[0].numel()
                v = torch.randn(p.shape, generator=g) / fan_in ** 0.5
            elif name.endswith("weight"):
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
            else:
                v = 0.1 * torch.randn(p.shape, generator=g)
            p.copy_(v.to(p.device, p.dtype))
    return module


def parameter_checksum(module) -> float:
    return float(sum(p.detach().double().sum().item() for p in module.parameters()))
