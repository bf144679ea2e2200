from .policies import BMSY, escapement, msy  # noqa: F401
