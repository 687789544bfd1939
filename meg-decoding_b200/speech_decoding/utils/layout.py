from meg_decoding.utils.layout import ch_locations_2d  # noqa: F401
