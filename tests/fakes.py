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
