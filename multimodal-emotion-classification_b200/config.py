"""Mirror of the audio constants of the reference's config.py:57-59 (the only Config fields the
speech path reads; bound as default arguments at import time, audio_preprocessing.py:12,22)."""


class Config:
    SAMPLE_RATE = 22050
    AUDIO_DURATION = 3
    N_MFCC = 40
