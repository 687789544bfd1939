from meg_decoding.utils.loss import *  # noqa: F401,F403
from meg_decoding.utils.loss import CLIPLoss  # noqa: F401
