"""Case definitions shared by ``make_golden.py`` (which runs the real reference) and the tests.

Pure data: JSON-style configs exactly as a user of the reference would write them.
"""

README_FBANK = {
    "name": "stft",
    "bank": "fbank",
    "frame_length_ms": 25,
    "include_energy": True,
    "pad_to_nearest_power_of_two": True,
    "window_function": "hanning",
    "use_power": True,
}

KALDI_FBANK = {  # tests/data/fbank.json of the reference
    "name": "stft",
    "bank": {
        "name": "fbank",
        "num_filts": 40,
        "low_hz": 20,
        "high_hz": 8000,
        "sampling_rate": 16000,
        "analytic": False,
    },
    "frame_length_ms": 25,
    "frame_shift_ms": 10,
    "frame_style": "centered",
    "include_energy": False,
    "pad_to_nearest_power_of_two": True,
    "window_function": "hanning",
    "use_log": True,
    "use_power": True,
    "kaldi_shift": True,
}

GAMMATONE_64 = {  # BASELINE config 3
    "name": "stft",
    "bank": {"name": "gammatone", "scaling_function": "mel", "num_filts": 64, "erb": True},
    "frame_length_ms": 25,
    "use_power": True,
}

# name -> (computer config, signal spec); signal spec = ("randn", seed, length) or ("wav", n)
STFT_CASES = {
    "readme_fbank_wav": (README_FBANK, ("wav", 48000)),
    "readme_fbank_noise": (README_FBANK, ("randn", 1, 16000)),
    "kaldi_fbank": (KALDI_FBANK, ("randn", 2, 8000)),
    "gammatone64": (GAMMATONE_64, ("randn", 3, 16000)),
    "gammatone64_L512": (dict(GAMMATONE_64, frame_length_ms=32), ("randn", 4, 8000)),
    "gabor41_power": (
        {
            "name": "stft",
            "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41},
            "frame_length_ms": 25,
            "use_power": True,
            "include_energy": True,
        },
        ("randn", 5, 8000),
    ),
    "gabor40_nopad_mag": (
        {
            "name": "stft",
            "bank": {"name": "gabor", "scaling_function": "mel"},
            "frame_length_ms": 25,
            "pad_to_nearest_power_of_two": False,
            "use_power": False,
        },
        ("randn", 6, 4000),
    ),
    "tri_bark_analytic": (
        {
            "name": "stft",
            "bank": {"name": "tri", "scaling_function": "bark", "analytic": True, "num_filts": 23},
            "frame_length_ms": 20,
            "frame_shift_ms": 5,
            "window_function": "hamming",
            "use_power": False,
            "use_log": False,
            "include_energy": True,
        },
        ("randn", 7, 6000),
    ),
    "fbank_causal_gamma_8k": (
        {
            "name": "stft",
            "bank": {"name": "fbank", "num_filts": 24, "sampling_rate": 8000, "high_hz": 3800},
            "frame_length_ms": 25,
            "frame_style": "causal",
            "use_power": True,
            "include_energy": True,
        },
        ("randn", 8, 4000),
    ),
    "fbank_blackman_1024": (
        {
            "name": "stft",
            "bank": {"name": "fbank", "num_filts": 40},
            "frame_length_ms": 50,
            "frame_shift_ms": 12.5,
            "window_function": "blackman",
            "use_power": True,
            "kaldi_shift": True,
        },
        ("randn", 9, 8000),
    ),
    "tri_octave_odd_shift": (  # odd frame shift -> generic (direct DFT) kernel
        {
            "name": "stft",
            "bank": {"name": "tri", "scaling_function": {"name": "octave", "low_hz": 20}, "num_filts": 10},
            "frame_length_ms": 25,
            "frame_shift_ms": 9.9375,
            "window_function": "bartlett",
            "use_power": True,
        },
        ("randn", 10, 3000),
    ),
}

# the edge lengths SURVEY.md 8(d) asks for, run with the README config on one seeded signal
EDGE_LENGTHS = (0, 1, 200, 201, 240, 399, 400, 401, 559, 560, 561, 1000)

SI_CASES = {
    "si_gabor41": (
        {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41}},
        ("randn", 11, 4000),
    ),
    "si_gabor_energy_power": (
        {
            "name": "si",
            "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 12},
            "include_energy": True,
            "use_power": True,
        },
        ("randn", 12, 2500),
    ),
    "si_gammatone_causal": (
        {
            "name": "si",
            "bank": {"name": "gammatone", "scaling_function": "mel", "num_filts": 40},
            "frame_shift_ms": 10,
        },
        ("randn", 13, 3000),
    ),
    "si_fbank_real_nolog": (
        {
            "name": "si",
            "bank": {"name": "tri", "scaling_function": "mel", "num_filts": 8, "sampling_rate": 8000},
            "use_log": False,
            "frame_shift_ms": 20,
        },
        ("randn", 14, 2400),
    ),
}

# supports beyond one 1024-point block (``si_long.npz``): the kernels with 1024 * R-point blocks
SI_LONG_CASES = {
    "si_fbank40": ({"name": "si", "bank": "fbank"}, ("randn", 21, 20000)),  # 6 987 taps, real filters, R = 16
    "si_gabor128_power": (  # 838 taps, R = 2
        {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 128}, "use_power": True},
        ("randn", 22, 9000),
    ),
    "si_fbank8_energy": (  # 2 206 taps, R = 4, energy column (a pure delay as the first filter)
        {"name": "si", "bank": {"name": "fbank", "num_filts": 8}, "include_energy": True},
        ("randn", 23, 12000),
    ),
    "si_fbank16_power_nolog": (  # 3 671 taps, R = 8
        {"name": "si", "bank": {"name": "fbank", "num_filts": 16}, "use_power": True, "use_log": False},
        ("randn", 24, 15000),
    ),
    "si_gammatone100_causal": (  # 934 taps, causal pooling window, complex filters, R = 2
        {"name": "si", "bank": {"name": "gammatone", "scaling_function": "mel", "num_filts": 100}},
        ("randn", 25, 7000),
    ),
}

BANK_CASES = {  # table-level goldens: supports, truncated / frequency / impulse responses
    "fbank40": ({"name": "fbank"}, 512),
    "fbank24_8k": ({"name": "fbank", "num_filts": 24, "sampling_rate": 8000, "high_hz": 3800}, 256),
    "tri_bark": ({"name": "tri", "scaling_function": "bark", "num_filts": 11, "sampling_rate": 8000}, 200),
    "tri_mel_analytic": ({"name": "tri", "scaling_function": "mel", "num_filts": 11, "analytic": True}, 256),
    "gabor_mel": ({"name": "gabor", "scaling_function": "mel", "num_filts": 11, "sampling_rate": 8000}, 256),
    "gabor_l2_erb": (
        {"name": "gabor", "scaling_function": "bark", "num_filts": 7, "scale_l2_norm": True, "erb": True},
        300,
    ),
    "gammatone_mel": ({"name": "gammatone", "scaling_function": "mel", "num_filts": 11, "sampling_rate": 8000}, 256),
    "gammatone_centered_l2": (
        {
            "name": "gammatone",
            "scaling_function": {"name": "linear", "low_hz": 0, "slope_hz": 0.5},
            "num_filts": 6,
            "max_centered": True,
            "scale_l2_norm": True,
            "order": 2,
        },
        512,
    ),
}

WINDOW_CASES = (
    ("hanning", 400),
    ("hamming", 255),
    ("blackman", 512),
    ("bartlett", 33),
    ("gamma", 400),
    ({"name": "gamma", "order": 1, "peak": 0.5}, 100),
    ("hann", 1),
)

# ---- round 2 (extra.npz) ----------------------------------------------------------------------
SI_GABOR_41 = {"name": "si", "bank": {"name": "gabor", "scaling_function": "mel", "num_filts": 41}}  # BASELINE config 4
C4_SEED, C4_SAMPLES = 77, 60 * 16000  # one 60 s utterance

# post.Stack: (name, constructor kwargs, axis passed to apply)
STACK_CASES = [
    ("k3", dict(num_vectors=3), -1),
    ("k4_edge", dict(num_vectors=4, pad_mode="edge"), -1),
    ("k5_const", dict(num_vectors=5, pad_mode="constant", constant_values=2.0), 1),
    ("k1", dict(num_vectors=1), -1),
    ("k2_time1", dict(num_vectors=2, time_axis=1), 0),
]
