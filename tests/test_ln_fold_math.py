"""Algebra behind the LayerNorm-folded GEMM epilogues (gemm_tc.cuh), checked in float64 on the CPU:
    Linear(LayerNorm(x))[n] = rstd * sum_k x[k] W'[n,k] + b'[n]
with W' = W diag(gamma) whose rows are centred and b' = b + W beta, and the two-term fp16 split of the residual stream
(hi = fp16(x), lo = fp16(x - hi)) that the engine stores instead of fp32."""
import torch


def test_centred_fold_equals_layernorm_linear():
    g = torch.Generator().manual_seed(0)
    M, D, N = 37, 256, 96
    x = torch.randn(M, D, generator=g, dtype=torch.float64) * 2.0 + 0.8      # non-zero row mean
    x[:, 3] += 50.0                                                           # a massive channel
    gamma = torch.rand(D, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(D, generator=g, dtype=torch.float64) * 0.3
    W = torch.randn(N, D, generator=g, dtype=torch.float64) * 0.1
    b = torch.randn(N, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-6), W, b)

    Wf = W * gamma
    Wf = Wf - Wf.mean(1, keepdim=True)                    # fold_ln_linear_kernel
    b2 = b + W @ beta
    mean = x.mean(-1, keepdim=True)                       # from the (sum, sum of squares) partials of the epilogue
    var = (x * x).mean(-1, keepdim=True) - mean * mean
    rstd = torch.rsqrt(var + 1e-6)
    got = rstd * (x @ Wf.t()) + b2                        # EPI_LN_BIAS_F16: no mean term left
    assert (got - ref).abs().max().item() < 1e-10


def test_two_term_fp16_split_keeps_22_bits():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4096, generator=g) * 3.0
    x[::97] *= 200.0                                      # massive activations, still far below the fp16 maximum
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    rec = hi.float() + lo.float()
    # 11 + 11 significant bits, down to the fp16 subnormal spacing 2^-24 of the lo plane (absolute floor 3e-8: the
    # residual stream is O(1) with a LayerNorm behind it, far above that)
    bound = torch.maximum(x.abs() * 2.0 ** -21, torch.tensor(2.0 ** -24))
    assert bool(((rec - x).abs() <= bound).all()), ((rec - x).abs() / bound).max().item()
    assert torch.equal(hi, x.to(torch.float16))           # the hi plane IS the tensor-core operand
