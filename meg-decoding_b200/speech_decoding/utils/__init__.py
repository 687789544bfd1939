from meg_decoding import _extend_path_with_reference

_extend_path_with_reference(__path__, "meg_decoding", "utils")
