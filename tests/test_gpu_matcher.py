"""-m gpu parity tests of the matcher kernels (through the C ABI) vs the CPU oracle and the golden vector."""
import pytest
import torch

from nerf_loc_b200 import params, synthetic as syn
from nerf_loc_b200.config import default_args
from nerf_loc_b200.matcher import Matcher
from oracle import matcher_oracle as MO
from oracle.make_golden import matcher_inputs
from tests.common import golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-4


def cuda_matcher(seed, planted=False):
    sd = syn.synthetic_state_dict(params.matcher_shapes(), seed)
    if planted:
        # a scoring MLP that is monotone in the descriptor correlation: planted pairs get well separated scores
        for k in ("coarse_matcher.mlps.0.weight", "coarse_matcher.mlps.2.weight", "coarse_matcher.mlps.4.weight"):
            sd[k] = sd[k].abs()
        for k in ("coarse_matcher.mlps.0.bias", "coarse_matcher.mlps.2.bias", "coarse_matcher.mlps.4.bias"):
            sd[k] = torch.full_like(sd[k], -3.0 if k.endswith("4.bias") else 0.0)
        sd["coarse_matcher.mlps.4.weight"] = sd["coarse_matcher.mlps.4.weight"] * 0.03
    m = Matcher(default_args(), 192, 192, 192).eval()
    m.load_state_dict(sd)
    return m.cuda(), sd


def to_cuda(d):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}


def test_matcher_forward_vs_golden():
    m, sd = cuda_matcher(99)
    g = golden("matcher_small")
    out = m(to_cuda(matcher_inputs()))
    assert relerr(out["score_matrix"].cpu(), g["score_matrix"]) < TOL
    assert torch.equal(out["i_ids"].cpu(), g["i_ids"]) and torch.equal(out["j_ids"].cpu(), g["j_ids"])
    for k in ("expec_f", "mkps2d_f", "mkps2d_c"):
        assert relerr(out[k].cpu(), g[k]) < TOL, k
    assert out["i_ids"].dtype == torch.int64 and out["pairs"][0] is out["i_ids"]


@pytest.mark.parametrize("N,M", [(1, 1), (37, 300), (513, 1111)])
def test_s2d_scores_ragged_sizes(N, M):
    m, sd = cuda_matcher(3)
    g = torch.Generator().manual_seed(N * 7 + M)
    d0, d1 = torch.randn(N, 192, generator=g), torch.randn(M, 192, generator=g)
    with torch.no_grad():
        ref = MO.s2d_scores(sd, "coarse_matcher", d0, d1)
    score, i_ids, j_ids = m.s2d(d0.cuda(), d1.cuda(), 0.2)
    assert relerr(score.cpu(), ref) < TOL
    # the match rule applied to OUR scores is exact (integer outputs)
    ri, rj = MO.mutual_matches(score.cpu(), 0.2)
    assert torch.equal(i_ids.cpu(), ri) and torch.equal(j_ids.cpu(), rj)


def test_planted_matches_are_recovered_identically():
    m, sd = cuda_matcher(5, planted=True)
    g = torch.Generator().manual_seed(11)
    N, M, P = 512, 1200, 256
    d1 = torch.randn(M, 192, generator=g)
    d0 = torch.randn(N, 192, generator=g)
    cells = torch.randperm(M, generator=g)[:P]
    d0[:P] = d1[cells] + 0.05 * torch.randn(P, 192, generator=g)
    with torch.no_grad():
        ref = MO.s2d_scores(sd, "coarse_matcher", d0, d1)
    ri, rj = MO.mutual_matches(ref, 0.2)
    score, i_ids, j_ids = m.s2d(d0.cuda(), d1.cuda(), 0.2)
    assert relerr(score.cpu(), ref) < TOL
    assert torch.equal(i_ids.cpu(), ri) and torch.equal(j_ids.cpu(), rj)
    assert len(ri) >= P and torch.equal(rj[:P], cells)


def test_mutual_rule_ties_and_empty():
    m, _ = cuda_matcher(1)
    L = __import__("nerf_loc_b200._lib", fromlist=["load"])
    lib = L.load()
    s = torch.tensor([[0.9, 0.9, 0.1], [0.1, 0.15, 0.1], [0.3, 0.2, 0.8], [0.3, 0.1, 0.8]])
    ri, rj = MO.mutual_matches(s, 0.2)
    sc = s.cuda()
    N, M = s.shape
    i_ids = torch.empty(N, dtype=torch.int64, device="cuda")
    j_ids = torch.empty(N, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    nb = lib.nlb_mutual_scratch_bytes(N, M)
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    L.check(lib.nlb_mutual_matches(L.ptr(sc), N, M, 0.2, L.ptr(i_ids), L.ptr(j_ids), L.ptr(cnt), L.ptr(scratch), nb, L.stream()))
    n = int(cnt.item())
    assert torch.equal(i_ids[:n].cpu(), ri) and torch.equal(j_ids[:n].cpu(), rj)
    L.check(lib.nlb_mutual_matches(L.ptr(sc), N, M, 0.95, L.ptr(i_ids), L.ptr(j_ids), L.ptr(cnt), L.ptr(scratch), nb, L.stream()))
    assert int(cnt.item()) == 0


def test_fine_stage_vs_oracle():
    m, sd = cuda_matcher(8)
    g = torch.Generator().manual_seed(2)
    hc, wc = 9, 11
    feat = torch.randn(1, hc * 2, wc * 2, 192, generator=g)
    j = torch.tensor([0, 10, 5, wc * hc - 1, 37, 37, 60])
    with torch.no_grad():
        ref_w = torch.nn.functional.linear(MO.fine_windows(feat.permute(0, 3, 1, 2), j, 2), sd["fine_preprocess.proj.weight"],
                                           sd["fine_preprocess.proj.bias"])
    wins = m.fine_windows(feat[0].cuda(), j.cuda(), 2, wc)
    assert relerr(wins.cpu(), ref_w) < TOL
    f0 = torch.randn(len(j), 192, generator=g)
    mk = torch.rand(len(j), 2, generator=g) * 10
    with torch.no_grad():
        e_ref, k_ref = MO.fine_match(sd, "fine_matcher", f0, ref_w, mk)
    e, k = m.fine_match(f0.cuda(), wins, mk.cuda())
    assert relerr(e.cpu(), e_ref) < TOL and relerr(k.cpu(), k_ref) < TOL


def test_empty_descriptor_sets_are_refused():
    m, _ = cuda_matcher(1)
    with pytest.raises(AssertionError):
        m.coarse_matcher(torch.empty(0, 192, device="cuda"), torch.randn(4, 192, device="cuda"), {})
