"""Stand-ins shared by the wrapper-stack tests (the same arithmetic as scripts/make_golden.py:FakeDiffusersUNet,
which produced tests/golden/wrappers.pt from the REFERENCE's own wrapper classes)."""
import torch


def fake_unet_math(latents, t, encoder_hidden_states):
    x = latents[:, :4].float()
    tt = torch.as_tensor(t).float().reshape(-1).to(latents.device)
    tt = tt.expand(latents.shape[0]) if tt.numel() == 1 else tt
    e = encoder_hidden_states.float().mean(dim=(1, 2))
    out = 0.7 * torch.tanh(x) + 0.001 * tt[:, None, None, None] * x.roll(1, -1) + 0.1 * e[:, None, None, None]
    if latents.shape[1] > 4:
        w = torch.arange(1, latents.shape[1] - 3, dtype=torch.float32, device=latents.device)[None, :, None, None]
        out = out + 0.05 * (latents[:, 4:].float() * w).sum(dim=1, keepdim=True)
    return out.to(latents.dtype)


class FakeDiffusersUNet:
    class _Out:
        def __init__(self, sample):
            self.sample = sample

    def __call__(self, latents, t, *, encoder_hidden_states, **kwargs):
        return self._Out(fake_unet_math(latents, t, encoder_hidden_states))


# ---- hint wrappers (scripts/make_golden.py:pin_hints ran the REFERENCE's UNetWithControlnet / UNetWithT2I over these)
from types import SimpleNamespace  # noqa: E402

HINT_SHAPES = [(2, 8, 8, 8), (2, 8, 8, 8), (2, 16, 4, 4)]


class FakeHintControlnet:
    def __init__(self, seed, cfg_only):
        self.w = [torch.randn(s[1:], generator=torch.Generator().manual_seed(seed + i)) for i, s in enumerate(HINT_SHAPES)]
        self.cfg_only = cfg_only

    def __call__(self, latents, t, encoder_hidden_states, cfg_meta=None):
        if self.cfg_only and cfg_meta == "u":
            return SimpleNamespace(down_block_res_samples=[torch.tensor(0)] * 2, mid_block_res_sample=torch.tensor(0))
        k = latents.mean(dim=(1, 2, 3)) + 0.01 * torch.as_tensor(t).float().reshape(-1) + encoder_hidden_states.mean(dim=(1, 2))
        res = [k[:, None, None, None] * w for w in self.w]
        if self.cfg_only and cfg_meta == "f":
            res = [torch.cat([torch.zeros_like(r.chunk(2)[0]), r.chunk(2)[1]]) for r in res]
        return SimpleNamespace(down_block_res_samples=res[:2], mid_block_res_sample=res[2])

class FakeHintAdapter:
    def __init__(self, seed, cfg_only):
        self.state = [torch.randn(1, 4 * (i + 1), 8 >> i, 8 >> i, generator=torch.Generator().manual_seed(seed + i)) for i in range(4)]
        self.cfg_only = cfg_only
        self.fuser = None

    def coadapter_type(self):
        return False

    def __call__(self):
        return self.state

class FakeHintUNet:
    def __call__(self, latents, t, **kw):
        out = latents * 0.5 + 0.001 * torch.as_tensor(t).float().reshape(-1)[:, None, None, None]
        out = out + kw["encoder_hidden_states"].mean(dim=(1, 2))[:, None, None, None]
        for r in kw.get("down_block_additional_residuals") or []:
            out = out + r.mean(dim=(1, 2, 3), keepdim=True)
        if kw.get("mid_block_additional_residual") is not None:
            out = out + 2 * kw["mid_block_additional_residual"].mean(dim=(1, 2, 3), keepdim=True)
        for i, a in enumerate(kw.get("adapter_states") or []):
            out = out + (i + 1) * a.mean(dim=(1, 2, 3), keepdim=True)
        return out



class FakeHintStyleAdapter:
    """A style adapter as UNetWithT2I sees it: its state is a tensor of context tokens [1, T, C], not a list."""

    def __init__(self, seed, tokens=3, dim=6):
        self.state = torch.randn(1, tokens, dim, generator=torch.Generator().manual_seed(seed))
        self.cfg_only = True
        self.fuser = None

    def coadapter_type(self):
        return False

    def __call__(self):
        return self.state
