"""Import shims that let the reference's agent module load without langchain / langgraph / copilotkit.

The reference (agent/game_agent_v2.py:14-46) imports five packages that are not installed here.  Only a thin
slice of each is used by the three hot-path nodes:

* `langchain.tools.tool`            - decorator; the nodes only read `.name` and pass the tool to bind_tools
* `langchain.chat_models.init_chat_model` - returns the chat model; here: the stub installed with set_model()
* `langchain_core.messages.*`       - plain message classes (content / tool_calls / tool_call_id)
* `langchain_core.runnables.RunnableConfig`, `langgraph.types.Command`, `langgraph.graph.{StateGraph, END}`
* `copilotkit.CopilotKitState`      - base class of AgentState; a dict is enough for `state.get(...)`

The reference also opens a log file under /home/lee at import (game_agent_v2.py:76-83); FileHandler and
makedirs are neutralised for the duration of the import.  Nothing in the reference is modified.
"""
from __future__ import annotations

import logging
import os
import sys
import types
from dataclasses import dataclass, field
from typing import Any, Dict

REFERENCE_ROOT = "/root/reference"
_model_holder = {"model": None}


def set_model(model) -> None:
    _model_holder["model"] = model


class _Msg:
    def __init__(self, content: Any = "", tool_calls=None, tool_call_id=None, **kw):
        self.content = content
        self.tool_calls = list(tool_calls or [])
        self.tool_call_id = tool_call_id
        self.additional_kwargs = kw


class SystemMessage(_Msg):
    pass


class HumanMessage(_Msg):
    pass


class AIMessage(_Msg):
    pass


class ToolMessage(_Msg):
    pass


@dataclass
class Command:
    goto: Any = None
    update: Dict[str, Any] = field(default_factory=dict)

    def __class_getitem__(cls, item):        # Command[Literal[...]] in annotations
        return cls


class _StateGraph:
    def __init__(self, *a, **k):
        self.nodes = {}

    def add_node(self, name, fn):
        self.nodes[name] = fn

    def __getattr__(self, name):
        return lambda *a, **k: self


def _tool(fn):
    fn.name = fn.__name__
    return fn


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "agent"))


_loaded = {}


def load_reference():
    """Returns the reference's game_agent_v2 module (imported once)."""
    if "mod" in _loaded:
        return _loaded["mod"]
    if not available():
        raise RuntimeError("Oracle A needs the reference checkout at %s" % REFERENCE_ROOT)
    _mod("langchain")
    _mod("langchain.tools", tool=_tool)
    _mod("langchain.chat_models", init_chat_model=lambda *a, **k: _model_holder["model"])
    _mod("langchain_core")
    _mod("langchain_core.messages", SystemMessage=SystemMessage, HumanMessage=HumanMessage, AIMessage=AIMessage,
         ToolMessage=ToolMessage, BaseMessage=_Msg)
    _mod("langchain_core.runnables", RunnableConfig=dict)
    _mod("langgraph")
    _mod("langgraph.types", Command=Command)
    _mod("langgraph.graph", StateGraph=_StateGraph, END="__end__")
    _mod("copilotkit", CopilotKitState=dict)
    agent_dir = os.path.join(REFERENCE_ROOT, "agent")
    if agent_dir not in sys.path:
        sys.path.insert(0, agent_dir)
    real_fh, real_makedirs = logging.FileHandler, os.makedirs
    logging.FileHandler = lambda *a, **k: logging.NullHandler()       # no /home/lee/... log file
    os.makedirs = lambda *a, **k: None
    try:
        import game_agent_v2 as mod                                   # the UNMODIFIED reference module
    finally:
        logging.FileHandler, os.makedirs = real_fh, real_makedirs
    mod.logger.handlers.clear()
    mod.logger.setLevel(logging.CRITICAL)
    logging.getLogger("tools.utils").setLevel(logging.CRITICAL)
    _loaded["mod"] = mod
    return mod
