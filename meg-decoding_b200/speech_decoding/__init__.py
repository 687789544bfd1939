"""Stale package name still imported by the reference's ``train.py:15-25`` (renamed ``meg_decoding`` upstream)."""
from meg_decoding import _extend_path_with_reference

# the renamed package of the reference checkout backs everything the alias does not provide (dataclass, utils.*)
_extend_path_with_reference(__path__, "meg_decoding")
