class EasyDict(dict):
    """Attribute-access dict (the subset of `easydict` config/config.py uses)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value
