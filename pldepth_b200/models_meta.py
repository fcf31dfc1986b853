"""String-keyed parameter bag the samplers read (mirror of pldepth/models/models_meta.py:27-70).

Any object with ``get_parameter(name, default=None)`` works; this
class is provided so the drop-in can be used without the reference installed.
"""
import copy


class ModelParameters(object):
    def __init__(self, **initial):
        self.parameters = dict(initial)

    def set_parameter(self, name, value):
        self.parameters[name] = value

    def get_parameter(self, name, default=None):
        return self.parameters.get(name, default)

    def get_parameter_string(self):
        return "_".join("%s_%s" % (k, v) for k, v in self.parameters.items())

    def duplicate(self):
        other = ModelParameters()
        other.parameters = copy.deepcopy(self.parameters)
        return other
