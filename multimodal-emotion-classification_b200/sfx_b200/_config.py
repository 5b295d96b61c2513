"""The Config the speech path reads (reference config.py:52-53,57-59).  The reference's own `config` module wins when it is
importable (drop-in use: PYTHONPATH=multimodal-emotion-classification_b200:/path/to/reference); the 5-attribute mirror
sfx_b200/config.py is the fallback for standalone use.  This package never puts a module called `config`, `inference` or
`model_training` on the path: only `preprocessing` shadows the reference."""


def resolve():
    try:
        from config import Config  # the reference's
        if all(hasattr(Config, k) for k in ("EMOTIONS", "SAMPLE_RATE", "AUDIO_DURATION", "N_MFCC")):
            return Config
    except Exception:      # noqa: BLE001  (absent, or its own imports fail)
        pass
    from .config import Config
    return Config


Config = resolve()
