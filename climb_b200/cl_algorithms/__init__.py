from .adapters import AdapterHandler
from .ewc import EWC
from .experience_replay import ExperienceReplayMemory
