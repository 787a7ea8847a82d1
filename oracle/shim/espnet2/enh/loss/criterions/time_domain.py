"""Oracle shim (TEST INFRASTRUCTURE): restatement of espnet2 time-domain criteria used by the reference
(baseline_code/d_model.py:24-25,74,80; baseline_code/flow_model.py:22).  SURVEY.md Appendix A / §8a row a15.
PARITY UNPINNED for these two losses: no espnet install is available to cross-check get_magnitude()."""
import torch
from espnet2.enh.encoder.stft_encoder import STFTEncoder


class SISNRLoss(torch.nn.Module):
    def __init__(self, clamp_db=None, zero_mean=True, eps=None):
        super().__init__()
        self.clamp_db, self.zero_mean = clamp_db, zero_mean
        self.eps = 1e-8 if eps is None else eps

    def forward(self, ref, est):
        assert ref.shape == est.shape
        if self.zero_mean:
            ref = ref - ref.mean(dim=-1, keepdim=True)
            est = est - est.mean(dim=-1, keepdim=True)
        energy = torch.sum(ref ** 2, dim=-1, keepdim=True) + self.eps
        proj = torch.sum(ref * est, dim=-1, keepdim=True) * ref / energy
        noise = est - proj
        ratio = torch.sum(proj ** 2, dim=-1) / (torch.sum(noise ** 2, dim=-1) + self.eps)
        return -10 * torch.log10(ratio + self.eps)


class MultiResL1SpecLoss(torch.nn.Module):
    def __init__(self, window_sz=(512,), hop_sz=None, eps=1e-8, time_domain_weight=0.5,
                 normalize_variance=False, reduction="sum"):
        super().__init__()
        assert all(w % 2 == 0 for w in window_sz)
        self.window_sz = list(window_sz)
        self.hop_sz = [w // 2 for w in window_sz] if hop_sz is None else list(hop_sz)
        self.eps, self.time_domain_weight = eps, time_domain_weight
        self.normalize_variance, self.reduction = normalize_variance, reduction
        self.stft_encoders = torch.nn.ModuleList(
            STFTEncoder(n_fft=w, win_length=w, hop_length=h, window=None, center=True,
                        normalized=False, onesided=True) for w, h in zip(self.window_sz, self.hop_sz))

    def forward(self, target, estimate):
        if self.normalize_variance:
            target = target / torch.std(target, dim=1, keepdim=True)
            estimate = estimate / torch.std(estimate, dim=1, keepdim=True)
        alpha = torch.sum(estimate * target, -1, keepdim=True) / (torch.sum(estimate ** 2, -1, keepdim=True) + self.eps)
        red = torch.sum if self.reduction == "sum" else torch.mean
        td = red((estimate * alpha - target).abs(), dim=-1)
        if not len(self.stft_encoders):
            return td
        lens = torch.full((target.size(0),), target.size(1), dtype=torch.long, device=target.device)
        sp = torch.zeros_like(td)
        for enc in self.stft_encoders:
            tm = enc(target, lens)[0].abs()
            em = enc(estimate * alpha, lens)[0].abs()
            sp = sp + red((em - tm).abs(), dim=(1, 2))
        return td * self.time_domain_weight + (1 - self.time_domain_weight) * sp / len(self.stft_encoders)
