"""Deterministic stand-in "UNets" for sampler parity tests (TEST INFRASTRUCTURE).

Only correctly-rounded fp32 element-wise ops (mul/add/sub), so the result is bit-identical
on CPU and on the GPU and golden vectors made here stay valid on the B200."""
import torch


class ToyEps:
    """model(x, t, return_dict=False)[0] and model(x, t).sample, like diffusers' UNet2DModel
    (call conventions: iadb_bn.py:319, ddim_diffusers.py:679)."""

    def __init__(self, out_channels):
        self.out_channels = out_channels

    def _f(self, x, t):
        t = torch.as_tensor(t, device=x.device).float()
        t = t.reshape(-1, 1, 1, 1) if t.dim() else t
        C = x.shape[1]
        inp = x[:, :min(C, self.out_channels)] if self.out_channels < C else x
        heads = [inp * 0.5 - t]
        while sum(h.shape[1] for h in heads) < self.out_channels:
            heads.append(inp * -0.25 + t * t)
        return torch.cat(heads, 1)[:, :self.out_channels].contiguous()

    def __call__(self, x, t, return_dict=True):
        y = self._f(x, t)
        if return_dict:
            class _O:  # noqa: N801
                pass
            o = _O()
            o.sample = y
            return o
        return (y,)


class ToyCond(ToyEps):
    """For sample_iadb_conditional: input is cat([x, x_c], 1) (iadb_bn.py:406)."""

    def _f(self, xin, t):
        C = xin.shape[1] // 2
        x, xc = xin[:, :C], xin[:, C:]
        t = torch.as_tensor(t, device=xin.device).float().reshape(-1, 1, 1, 1)
        heads = [x * 0.5 - t + xc * 0.125]
        while sum(h.shape[1] for h in heads) < self.out_channels:
            heads.append(x * -0.25 + t * t)
        return torch.cat(heads, 1)[:, :self.out_channels].contiguous()
