"""Numeric constants of the SFC path (values are part of the contract, reference lib/constants.py)."""
INPUT_SAMPLE_RATE = 16_000      # Hz, model input
TARGET_SAMPLE_RATE = 49.95      # frames per second of the classifier output (not exactly 50)
WAV2VEC_FRAME_LEN = 20          # ms per wav2vec 2.0 frame, used for step <-> second conversions
HIDDEN_SIZE = 1024              # XLS-R-300m hidden width
NOISE_THRESHOLD = 0.1           # s
