from jumanji.environments.routing.connector.env import Connector  # noqa: F401
