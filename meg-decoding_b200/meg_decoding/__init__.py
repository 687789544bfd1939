"""Drop-in mirror of the reference's ``meg_decoding`` package for the training hot path (see models.py, utils/loss.py)."""
