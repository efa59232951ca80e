from .launcher import Launcher, SurrealDefaultLauncher  # noqa: F401
