from meg_decoding.models import *  # noqa: F401,F403
from meg_decoding.models import BrainEncoder, Classifier, get_model  # noqa: F401
