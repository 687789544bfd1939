"""Stale package name still imported by the reference's ``train.py:15-25`` (renamed ``meg_decoding`` upstream)."""
