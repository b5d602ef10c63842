class Error(Exception):
    pass


class DependencyNotInstalled(Error):
    pass


class UnregisteredEnv(Error):
    pass
