from .policies import BMSY, escapement, msy, user_action  # noqa: F401
