"""`jclip` API surface of the reference (jclip/__init__.py: `from .clip import *`)."""
from .clip import *  # noqa: F401,F403
from .clip import available_models, load, load_vlp, tokenize, load_state_dict  # noqa: F401
from .model import CLIP, build_model  # noqa: F401
