"""`CfgNode`: the subset of yacs.config.CfgNode the reference uses -- attribute access on a dict (so `json.dump(cfg)` works,
lib/train_recoder.py:24), `defrost()` / `freeze()` (train_stage2.py:192-199), `clone()` (config/stereo_human_config.py:56),
`merge_from_file()` (:60) with yacs' value decoding (YAML strings such as 'None' or '1e-5' are literal-evaluated) and its
rule that a file may only override keys that already exist, plus `merge_from_list`, `merge_from_other_cfg`, `dump`."""
import copy
from ast import literal_eval

import yaml


class CfgNode(dict):
    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # ---- attribute access ----
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    # ---- mutability ----
    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def _immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode.IMMUTABLE] = self.is_frozen()
        return out

    # ---- merging ----
    @staticmethod
    def _decode(value):
        if isinstance(value, dict):
            return CfgNode(value)
        if not isinstance(value, str):
            return value
        try:
            return literal_eval(value)
        except (ValueError, SyntaxError):
            return value

    def _merge(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError(f"Non-existent config key: {full}")
            v = self._decode(copy.deepcopy(v))
            if isinstance(self[k], CfgNode) and isinstance(v, dict):
                self[k]._merge(v, path + [k])
            else:
                self[k] = v

    def merge_from_other_cfg(self, other):
        self._merge(other, [])

    def merge_from_file(self, cfg_filename):
        with open(cfg_filename, "r") as f:
            loaded = yaml.safe_load(f) or {}
        self._merge(loaded, [])

    def merge_from_list(self, cfg_list):
        if len(cfg_list) % 2:
            raise ValueError("Override list has odd length")
        for full, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            *parents, leaf = full.split(".")
            for p in parents:
                node = node[p]
            if leaf not in node:
                raise KeyError(f"Non-existent config key: {full}")
            node[leaf] = self._decode(v)

    def dump(self, **kwargs):
        plain = lambda n: {k: plain(v) if isinstance(v, CfgNode) else v for k, v in n.items()}
        return yaml.safe_dump(plain(self), **kwargs)

    def __repr__(self):
        return f"CfgNode({dict.__repr__(self)})"
