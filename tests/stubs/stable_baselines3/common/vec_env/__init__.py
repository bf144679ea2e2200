from abc import ABC, abstractmethod


class VecEnv(ABC):
    """Abstract vectorised environment: the methods an SB3 algorithm calls on its `env`."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.sb3_base_initialised = True

    @abstractmethod
    def reset(self):
        ...

    @abstractmethod
    def step_async(self, actions):
        ...

    @abstractmethod
    def step_wait(self):
        ...

    @abstractmethod
    def close(self):
        ...

    @abstractmethod
    def get_attr(self, attr_name, indices=None):
        ...

    @abstractmethod
    def set_attr(self, attr_name, value, indices=None):
        ...

    @abstractmethod
    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        ...

    @abstractmethod
    def env_is_wrapped(self, wrapper_class, indices=None):
        ...

    @abstractmethod
    def seed(self, seed=None):
        ...

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    @property
    def unwrapped(self):
        return self
