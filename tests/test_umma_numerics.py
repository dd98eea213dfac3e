"""CPU model of the arithmetic of stft_umma_kernel (csrc/stft_umma.cuh): the 512-point spectrum of a windowed
frame from four 128-point real DFTs evaluated as fp16 two-term products accumulated in float32, followed by
the radix-4 step.  Pins the error budget quoted in DESIGN.md section 4.1b without a GPU: operand splits carry 22
bits, so a bin's error is about 2^-22 of the FRAME's peak bin, whatever the bin's own size."""
import numpy as np
import pytest


def split_fp16(x):
    """x = hi + lo with hi the 11-significant-bit rounding of x (exact in fp16) and lo the fp16 rounding of the rest"""
    bits = x.astype(np.float32).view(np.uint32)
    hi = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (x.astype(np.float32) - hi).astype(np.float16).astype(np.float32)
    assert np.array_equal(hi, hi.astype(np.float16).astype(np.float32))  # representable: nothing lost in the pack
    return hi, lo


def umma_power_spectrum(frame, window):
    L = len(frame)
    kch = 2 * ((L + 63) // 64)
    xw = np.zeros(32 * kch, dtype=np.float32)
    xw[:L] = frame.astype(np.float32) * window.astype(np.float32)
    peak = np.abs(xw).max()
    exponent = int(np.floor(np.log2(peak))) if peak > 0 else 0
    scale = np.float32(2.0 ** (14 - exponent))  # peak * scale in [2^14, 2^15)
    a_hi, a_lo = split_fp16(xw * scale)
    j = np.arange(8 * kch)
    n = np.arange(128)
    k1 = n % 64
    angle = 2 * np.pi * np.outer(j, k1) / 128.0
    basis = np.where(n[None, :] > 64, -np.sin(angle), np.cos(angle))
    basis[:, 64] = np.cos(np.pi * j)  # column 64 carries Re Y[64] in place of Im Y[0]
    b_hi = basis.astype(np.float16).astype(np.float32)
    b_lo = (basis - b_hi).astype(np.float16).astype(np.float32)
    power = np.zeros(257)
    y = []
    for r in range(4):
        ah, al = a_hi[r::4][: 8 * kch], a_lo[r::4][: 8 * kch]
        # float32 accumulation of the four products (the order inside the tensor core is not modelled)
        d = (ah @ b_hi + al @ b_hi + ah @ b_lo + al @ b_lo).astype(np.float32)
        y.append(d)
    y = np.array(y, dtype=np.float64)
    for k in range(65):
        if k == 0:
            yr = y[:, 0] + 0j
        elif k == 64:
            yr = y[:, 64] + 0j
        else:
            yr = y[:, k] + 1j * y[:, 64 + k]
        t = yr * np.exp(-2j * np.pi * np.arange(4) * k / 512.0)
        for k2 in range(4):
            x = np.sum(t * (-1j) ** (np.arange(4) * k2))
            bin_ = k + 128 * k2
            bin_ = 512 - bin_ if bin_ > 256 else bin_
            power[bin_] = abs(x) ** 2
    return power / float(scale) ** 2


@pytest.mark.parametrize("seed,amplitude", [(0, 1000.0), (1, 1e-3), (2, 30000.0)])
def test_fp16_two_term_dft_matches_rfft(seed, amplitude):
    rng = np.random.default_rng(seed)
    frame = rng.standard_normal(400) * amplitude
    window = np.hanning(400)
    want = np.abs(np.fft.rfft(frame.astype(np.float32).astype(np.float64) * window.astype(np.float32), 512)) ** 2
    got = umma_power_spectrum(frame, window)
    # amplitude error relative to the frame's largest bin: about 2^-22 (measured below 1e-6)
    err = np.abs(np.sqrt(got) - np.sqrt(want)).max() / np.sqrt(want.max())
    assert err < 1e-6
    # white noise: every bin is within 1e-4 of itself as well
    assert (np.abs(got - want) / want).max() < 1e-4


def test_bins_far_below_the_peak_feel_the_22_bits():
    """a tone 70 dB above the noise floor: the bins of the floor are only good to about 1e-3 relative -- the reason
    the speech golden sits at 8.9e-5 after the filter bank and why bf16 splits (16 bits) would not do"""
    rng = np.random.default_rng(5)
    n = np.arange(400)
    frame = 20000.0 * np.sin(2 * np.pi * 1000.0 / 16000.0 * n) + rng.standard_normal(400) * 2.0
    window = np.hanning(400)
    want = np.abs(np.fft.rfft(frame.astype(np.float32).astype(np.float64) * window.astype(np.float32), 512)) ** 2
    got = umma_power_spectrum(frame, window)
    err = np.abs(np.sqrt(got) - np.sqrt(want)).max() / np.sqrt(want.max())
    assert err < 1e-6
    floor_bins = want < 1e-6 * want.max()
    assert floor_bins.any() and (np.abs(got - want) / want)[floor_bins].max() < 0.05
