"""``from hugectr.inference import InferenceParams, CreateInferenceSession``."""
from hugectr_b200.inference import *  # noqa: F401,F403
from hugectr_b200.inference import CreateInferenceSession, InferenceParams  # noqa: F401
