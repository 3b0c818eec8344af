"""Minimal stand-in for the `yacs` package, only so that the UNMODIFIED reference config
(config/stereo_human_config.py:1 `from yacs.config import CfgNode as CN`) imports in an environment without yacs
(harness item).  Not part of the hot path."""
