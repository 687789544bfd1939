"""Drop-in mirror of the reference's ``meg_decoding`` package for the training hot path (see models.py, utils/loss.py)."""


def _extend_path_with_reference(path, *sub):
    """``MEG_DECODING_REF=/path/to/MEG-decoding``: sub-modules this drop-in does not provide (``dataclass``,
    ``utils.get_dataloaders``, ``utils.preproc_utils``, loggers, ...) resolve from the reference checkout, while the
    hot-path modules (``models``, ``utils.loss``, ``utils.layout``) keep resolving here because this directory comes
    first on the package's ``__path__`` -- so the reference's ``train.py`` runs unchanged with
    ``PYTHONPATH=<repo>/meg-decoding_b200:$MEG_DECODING_REF``."""
    import os
    ref = os.environ.get("MEG_DECODING_REF")
    if ref:
        p = os.path.join(ref, *sub)
        if os.path.isdir(p) and p not in path:
            path.append(p)


_extend_path_with_reference(__path__, "meg_decoding")
