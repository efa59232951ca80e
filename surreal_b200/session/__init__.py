from .config import Config, ConfigError, extend_config  # noqa: F401
from .default_configs import (BASE_LEARNER_CONFIG, BASE_ENV_CONFIG, BASE_SESSION_CONFIG,  # noqa: F401
                              LOCAL_SESSION_CONFIG)
