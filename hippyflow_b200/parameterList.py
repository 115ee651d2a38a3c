"""Minimal stand-in for ``hippylib.ParameterList``: a dict of ``name -> [value, description]`` whose
item access returns / sets the VALUE (the way hippyflow reads ``self.parameters['rank']``,
e.g. hippyflow/modeling/PODProjector.py:365)."""


class ParameterList:
    def __init__(self, data):
        self.data = data

    def __getitem__(self, key):
        if key not in self.data:
            raise ValueError(key)
        return self.data[key][0]

    def __setitem__(self, key, value):
        if key not in self.data:
            raise ValueError(key)
        self.data[key][0] = value

    def __contains__(self, key):
        return key in self.data

    def keys(self):
        return self.data.keys()

    def showMe(self, indent=""):
        for k in sorted(self.data.keys()):
            print(indent + "---")
            print(indent + k, "(default, description):", self.data[k])
