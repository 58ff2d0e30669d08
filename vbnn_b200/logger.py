"""logger.lua restated (reference logger.lua:1-47): one text file per metric id under `dir`, one value
per line -- the format visualize.py:25-31 (`read_data`: float(line.strip()) per line) polls.  The 14
per-layer diagnostics of VBLinear.lua:150-163 reach it through `VBLinear.update(opt)` when `opt.log`
is set and a global logger is installed with `init()` (the reference's global `Log`, main.lua:151)."""
from __future__ import annotations

import os

Log = None          # the reference's global `Log` (main.lua:151)


class Logger:
    def __init__(self, dir, append=False):                              # logger.lua:5-12
        os.makedirs(dir, exist_ok=True)
        self.dir = dir
        self.loggers = {}
        self._append = append

    def _create(self, id, mode):                                        # logger.lua:14-16
        self.loggers[id] = open(os.path.join(self.dir, id), mode)

    def add(self, id, value):                                           # logger.lua:18-26
        if id not in self.loggers:
            if not self._append:
                open(os.path.join(self.dir, id), "w").close()
            self._create(id, "a")
        self.loggers[id].write(_lua_number(value) + "\n")

    def append(self, id, value):                                        # logger.lua:28-34
        if id not in self.loggers or not self._append:
            if id in self.loggers:
                self.loggers[id].close()
            self._create(id, "a+")
            self._append = True
        self.add(id, value)

    def flush(self):                                                    # logger.lua:36-40
        for f in self.loggers.values():
            f.flush()

    def close(self):                                                    # logger.lua:42-46
        for f in self.loggers.values():
            f.close()
        self.loggers = {}


def _lua_number(v):
    """Lua's tostring(number) is "%.14g"."""
    return "%.14g" % float(v)


def init(dir, append=False):
    """Log = require('logger'):init(opt.network_name) (main.lua:151; append mode on resume, :148)."""
    global Log
    Log = Logger(dir, append)
    return Log


def read_data(filename):
    """visualize.py:25-31 (the consumer of the files), kept beside the writer for the round-trip test."""
    with open(filename) as f:
        try:
            return [float(x.strip()) for x in f.readlines()]
        except ValueError:
            return []
