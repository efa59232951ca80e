from .launcher import Launcher, SurrealDefaultLauncher, PipelinedEngine  # noqa: F401
