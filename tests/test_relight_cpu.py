"""Relighting pass, host-side parts (no kernels): Environment_Light tables / background lookup equal the reference's,
and the inverse-CDF sampler draws from the reference's multinomial distribution."""
import torch


def test_environment_light_tables(golden_rotated):
    from tensoir_b200.relight import Environment_Light
    fx = golden_rotated
    env = Environment_Light({"sunny": fx["env_rgb"].numpy()}, device='cpu')
    assert torch.equal(env.hdr_pdf_sample["sunny"], fx["env_pdf_sample"])
    assert torch.equal(env.hdr_pdf_return["sunny"], fx["env_pdf_return"])
    assert torch.equal(env.hdr_dir["sunny"], fx["env_dir"])
    assert torch.equal(env.get_light("sunny", fx["rays"][:, 3:]), fx["relight_bg_lookup"])
    # replaying the reference's multinomial indices reproduces its gathers exactly
    idx = fx["relight_idx"]
    ld, lr, lp = env.sample_light("sunny", idx.shape[0], idx.shape[1], light_dir_idx=idx)
    assert torch.equal(ld, fx["env_dir"].view(-1, 3)[idx])
    assert torch.equal(lp.squeeze(-1), fx["env_pdf_return"].view(-1)[idx])


def test_inverse_cdf_sampler_matches_pdf(golden_rotated):
    from tensoir_b200.relight import Environment_Light
    fx = golden_rotated
    env = Environment_Light({"sunny": fx["env_rgb"].numpy()}, device='cpu')
    torch.manual_seed(0)
    n = 400_000
    u = torch.rand(1, n, dtype=torch.float64)
    cdf = env._cdf["sunny"]
    idx = torch.searchsorted(cdf, u, right=True).clamp_(max=cdf.numel() - 1).reshape(-1)
    hist = torch.bincount(idx, minlength=cdf.numel()).double() / n
    pdf = fx["env_pdf_sample"].view(-1).double()
    # every bin within 5 sigma of its binomial expectation, and the sun bins (40x brighter) dominate
    sigma = torch.sqrt(pdf * (1 - pdf) / n)
    assert bool(((hist - pdf).abs() <= 5 * sigma + 1e-6).all())
    assert float(hist.max()) > 10 * float(hist.median())
