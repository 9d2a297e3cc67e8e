"""Simulation configuration (mirror of sailfish/config.py:17-91 LBConfig/LBConfigParser,
reduced to argparse: the rc-file lookup of the reference is control plane, out of scope)."""
import argparse


class LBConfig(argparse.Namespace):
    """sailfish/config.py:17-29."""
    @property
    def output_required(self):
        return bool(self.output) or self.mode == 'visualization'

    @property
    def needs_iteration_num(self):
        return self.time_dependence or self.access_pattern == 'AA'


class LBConfigParser(object):
    def __init__(self, description=None):
        self._parser = argparse.ArgumentParser(description=description)
        self.config = LBConfig()

    def add_group(self, name):
        return self._parser.add_argument_group(name)

    def set_defaults(self, defaults):
        return self._parser.set_defaults(**defaults)

    def parse(self, args, internal_defaults=None):
        config = LBConfig()
        if internal_defaults is not None:
            for k, v in internal_defaults.items():
                setattr(config, k, v)
        self._parser.parse_args(args, config)
        # internal, non-CLI switches (sailfish/config.py:78-90)
        for k, v in (('relaxation_enabled', True), ('propagation_enabled', True),
                     ('time_dependence', False), ('space_dependence', False), ('unit_test', False)):
            if not hasattr(config, k):
                setattr(config, k, v)
        self.config = config
        return config
