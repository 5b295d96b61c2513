"""Fallback mirror (used only when the reference's own config.py is not importable) of the constants of the reference's config.py the speech path reads: the audio settings (:57-59, bound as
default arguments at import time, audio_preprocessing.py:12,22) and the label list (:52-53, speech_inference.py:15)."""


class Config:
    EMOTIONS = ['happy', 'sad', 'angry', 'fear', 'disgust', 'surprise', 'neutral']   # reference config.py:52
    NUM_EMOTIONS = 7                                                                  # reference config.py:53
    SAMPLE_RATE = 22050
    AUDIO_DURATION = 3
    N_MFCC = 40
