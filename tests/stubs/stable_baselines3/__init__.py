"""Stand-in for stable-baselines3 (absent from the build image), just the piece FishingVecEnv binds to: the abstract VecEnv
base class with the abstract-method set of stable-baselines3 (0.9.0, the version the reference pins in
examples/requirements.txt:29, plus env_is_wrapped which 1.0 added).  TEST INFRASTRUCTURE ONLY: tests put this directory on
sys.path to check that the product class is a VecEnv subclass with nothing abstract left, the way `oracle/refstub` stands in for
gym.  The method list and signatures restate the public interface; no stable-baselines3 code is copied."""
__version__ = "0.0.0-stub"
