"""Stand-in for matplotlib so gym_fishing/envs/shared_env.py:3 imports; plotting is out of scope."""
