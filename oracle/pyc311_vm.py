"""A small interpreter for CPython 3.11 bytecode.  TEST INFRASTRUCTURE ONLY.

Why: upstream ships `models.bistride_ops` and the older `models.bsms_mgn` only as `*.cpython-311.pyc`.  This image
runs Python 3.12, which neither executes 3.11 bytecode nor even returns it unmodified (`code.co_code` re-encodes it
with 3.12's opcode tables).  To pin oracle/bistride_oracle.py against the reference ITSELF, this module (1) reads the
marshal stream of a 3.11 .pyc into plain `Code` records with the raw instruction bytes, and (2) executes those
instructions with ordinary Python objects (torch tensors, nn.Modules) -- enough of the 3.11 instruction set for the
straight-line numerical code in those two modules (calls, attribute access, arithmetic, subscripts, loops, list
comprehensions, class bodies, zero-argument super()).  No exception handling, generators or `with` blocks: the
modules do not use them on the paths exercised.

Used only by oracle/gen_bistride_golden.py (here, where /root/reference exists); nothing imports it at test time.
"""
from __future__ import annotations

import builtins
import importlib
import operator
import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

# ------------------------------------------------------------------------------------------------
# marshal reader (format version 4, the subset a .pyc uses)
# ------------------------------------------------------------------------------------------------
FLAG_REF = 0x80


@dataclass
class Code:
    argcount: int
    posonlyargcount: int
    kwonlyargcount: int
    stacksize: int
    flags: int
    code: bytes
    consts: tuple
    names: tuple
    localsplusnames: tuple
    localspluskinds: bytes
    filename: str
    name: str
    qualname: str
    firstlineno: int
    linetable: bytes
    exceptiontable: bytes


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.p, self.refs = data, 0, []

    def byte(self) -> int:
        b = self.d[self.p]
        self.p += 1
        return b

    def i32(self) -> int:
        v = struct.unpack_from("<i", self.d, self.p)[0]
        self.p += 4
        return v

    def take(self, n: int) -> bytes:
        b = self.d[self.p: self.p + n]
        self.p += n
        return b

    def obj(self) -> Any:
        t = self.byte()
        flag, t = t & FLAG_REF, chr(t & ~FLAG_REF)
        idx = None
        if flag:
            idx = len(self.refs)
            self.refs.append(None)
        v = self._payload(t)
        if idx is not None:
            self.refs[idx] = v
        return v

    def _payload(self, t: str) -> Any:
        if t == "0":
            raise ValueError("NULL object in marshal stream")
        if t == "N":
            return None
        if t == "F":
            return False
        if t == "T":
            return True
        if t == ".":
            return Ellipsis
        if t == "i":
            return self.i32()
        if t == "l":
            n = self.i32()
            digits = [struct.unpack_from("<H", self.take(2))[0] for _ in range(abs(n))]
            v = sum(dg << (15 * k) for k, dg in enumerate(digits))
            return -v if n < 0 else v
        if t == "g":
            return struct.unpack("<d", self.take(8))[0]
        if t == "y":
            return complex(*struct.unpack("<dd", self.take(16)))
        if t == "s":
            return self.take(self.i32())
        if t in "ut":
            return self.take(self.i32()).decode("utf-8", "surrogatepass")
        if t in "aA":
            return self.take(self.i32()).decode("latin-1")
        if t in "zZ":
            return self.take(self.byte()).decode("latin-1")
        if t == ")":
            n = self.byte()
            return tuple(self.obj() for _ in range(n))
        if t == "(":
            n = self.i32()
            return tuple(self.obj() for _ in range(n))
        if t == "[":
            n = self.i32()
            return [self.obj() for _ in range(n)]
        if t in "<>":
            n = self.i32()
            items = [self.obj() for _ in range(n)]
            return set(items) if t == "<" else frozenset(items)
        if t == "{":
            out = {}
            while True:
                if chr(self.d[self.p] & ~FLAG_REF) == "0":
                    self.p += 1
                    return out
                k = self.obj()
                out[k] = self.obj()
        if t == "r":
            return self.refs[self.i32()]
        if t == "c":
            a = [self.i32() for _ in range(5)]
            code = self.obj()
            consts, names, lpn, lpk = self.obj(), self.obj(), self.obj(), self.obj()
            filename, name, qualname = self.obj(), self.obj(), self.obj()
            first = self.i32()
            linetable, exctable = self.obj(), self.obj()
            return Code(*a, code, consts, names, lpn, lpk, filename, name, qualname, first, linetable, exctable)
        raise ValueError(f"unsupported marshal type {t!r} at {self.p - 1}")


def load_pyc(path: str) -> Code:
    data = open(path, "rb").read()
    magic = int.from_bytes(data[:2], "little")
    if not 3495 <= magic <= 3499:
        raise ValueError(f"{path}: magic {magic} is not CPython 3.11")
    return _Reader(data[16:]).obj()


# ------------------------------------------------------------------------------------------------
# CPython 3.11 opcodes (Lib/opcode.py of 3.11)
# ------------------------------------------------------------------------------------------------
OP = dict(
    CACHE=0, POP_TOP=1, PUSH_NULL=2, NOP=9, UNARY_POSITIVE=10, UNARY_NEGATIVE=11, UNARY_NOT=12, UNARY_INVERT=15,
    BINARY_SUBSCR=25, GET_LEN=30, STORE_SUBSCR=60, DELETE_SUBSCR=61, GET_ITER=68, LOAD_BUILD_CLASS=71,
    LOAD_ASSERTION_ERROR=74, LIST_TO_TUPLE=82, RETURN_VALUE=83, STORE_NAME=90, DELETE_NAME=91, UNPACK_SEQUENCE=92,
    FOR_ITER=93, STORE_ATTR=95, STORE_GLOBAL=97, SWAP=99, LOAD_CONST=100, LOAD_NAME=101, BUILD_TUPLE=102, BUILD_LIST=103,
    BUILD_SET=104, BUILD_MAP=105, LOAD_ATTR=106, COMPARE_OP=107, IMPORT_NAME=108, IMPORT_FROM=109, JUMP_FORWARD=110,
    JUMP_IF_FALSE_OR_POP=111, JUMP_IF_TRUE_OR_POP=112, POP_JUMP_FORWARD_IF_FALSE=114, POP_JUMP_FORWARD_IF_TRUE=115,
    LOAD_GLOBAL=116, IS_OP=117, CONTAINS_OP=118, COPY=120, BINARY_OP=122, LOAD_FAST=124, STORE_FAST=125, DELETE_FAST=126,
    POP_JUMP_FORWARD_IF_NOT_NONE=128, POP_JUMP_FORWARD_IF_NONE=129, RAISE_VARARGS=130, MAKE_FUNCTION=132, BUILD_SLICE=133,
    MAKE_CELL=135, LOAD_CLOSURE=136, LOAD_DEREF=137, STORE_DEREF=138, JUMP_BACKWARD=140, CALL_FUNCTION_EX=142,
    EXTENDED_ARG=144, LIST_APPEND=145, SET_ADD=146, MAP_ADD=147, LOAD_CLASSDEREF=148, COPY_FREE_VARS=149, RESUME=151,
    FORMAT_VALUE=155, BUILD_CONST_KEY_MAP=156, BUILD_STRING=157, LOAD_METHOD=160, LIST_EXTEND=162, SET_UPDATE=163,
    DICT_MERGE=164, DICT_UPDATE=165, PRECALL=166, CALL=171, KW_NAMES=172, POP_JUMP_BACKWARD_IF_NOT_NONE=173,
    POP_JUMP_BACKWARD_IF_NONE=174, POP_JUMP_BACKWARD_IF_FALSE=175, POP_JUMP_BACKWARD_IF_TRUE=176,
)
NAME = {v: k for k, v in OP.items()}
BINARY = {
    0: operator.add, 1: operator.and_, 2: operator.floordiv, 3: operator.lshift, 4: operator.matmul, 5: operator.mul,
    6: operator.mod, 7: operator.or_, 8: operator.pow, 9: operator.rshift, 10: operator.sub, 11: operator.truediv,
    12: operator.xor, 13: operator.iadd, 14: operator.iand, 15: operator.ifloordiv, 16: operator.ilshift,
    17: operator.imatmul, 18: operator.imul, 19: operator.imod, 20: operator.ior, 21: operator.ipow, 22: operator.irshift,
    23: operator.isub, 24: operator.itruediv, 25: operator.ixor,
}
COMPARE = {0: operator.lt, 1: operator.le, 2: operator.eq, 3: operator.ne, 4: operator.gt, 5: operator.ge}
CO_FAST_LOCAL, CO_FAST_CELL, CO_FAST_FREE = 0x20, 0x40, 0x80
_NULL = object()     # the NULL the 3.11 call protocol pushes under callables


class Cell:
    __slots__ = ("v",)

    def __init__(self, v=_NULL):
        self.v = v


class Function:
    """A function object whose body is 3.11 bytecode run by `run`."""

    def __init__(self, code: Code, globs: dict, defaults=(), kwdefaults=None, closure=()):
        self.code, self.globs, self.defaults, self.kwdefaults, self.closure = code, globs, defaults, kwdefaults or {}, closure
        self.__name__, self.__qualname__, self.__doc__ = code.name, code.qualname, (code.consts[0] if code.consts and isinstance(code.consts[0], str) else None)

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        return lambda *a, **k: self(obj, *a, **k)

    def __call__(self, *args, **kwargs):
        co = self.code
        names = co.localsplusnames
        nparams = co.argcount + co.kwonlyargcount
        if co.flags & 0x04 or co.flags & 0x08:
            raise NotImplementedError("*args / **kwargs parameters")
        if len(args) > co.argcount:
            raise TypeError(f"{co.name}() takes {co.argcount} positional arguments but {len(args)} were given")
        fast: List[Any] = [_NULL] * len(names)
        for i, a in enumerate(args):
            fast[i] = a
        for k, v in kwargs.items():
            if k not in names[:nparams]:
                raise TypeError(f"{co.name}() got an unexpected keyword argument {k!r}")
            i = names.index(k)
            if fast[i] is not _NULL:
                raise TypeError(f"{co.name}() got multiple values for argument {k!r}")
            fast[i] = v
        ndef = len(self.defaults)
        for i in range(co.argcount):
            if fast[i] is _NULL:
                j = i - (co.argcount - ndef)
                if j < 0:
                    raise TypeError(f"{co.name}() missing required argument {names[i]!r}")
                fast[i] = self.defaults[j]
        for i in range(co.argcount, nparams):
            if fast[i] is _NULL:
                fast[i] = self.kwdefaults[names[i]]
        return run(co, self.globs, fast, self.closure)


def _build_class(func: Function, name: str, *bases, **kw):
    ns: Dict[str, Any] = {}
    cell = Cell()
    run(func.code, func.globs, None, func.closure, class_ns=ns, class_cell=cell)
    ns.pop("__classcell__", None)
    meta = type(bases[0]) if bases else type
    cls = meta(name, bases, ns)
    cell.v = cls
    return cls


def run(co: Code, globs: dict, fast: Optional[list], closure: tuple = (), class_ns: Optional[dict] = None,
        class_cell: Optional[Cell] = None):
    code, consts, names = co.code, co.consts, co.names
    lpn, kinds = co.localsplusnames, co.localspluskinds
    if fast is None:
        fast = [_NULL] * len(lpn)
    # free variables live at the end of localsplus; COPY_FREE_VARS copies the closure there
    stack: List[Any] = []
    kwnames: Tuple[str, ...] = ()
    ip, ext = 0, 0
    builtin = builtins.__dict__
    first_arg = fast[0] if (co.argcount and fast) else None
    # the class body's implicit __class__ cell
    if class_cell is not None:
        for i, n in enumerate(lpn):
            if n == "__class__" and kinds[i] & CO_FAST_CELL:
                fast[i] = class_cell

    def lookup_global(n):
        if n in globs:
            return globs[n]
        if n in builtin:
            return builtin[n]
        raise NameError(n)

    while True:
        op, arg = code[ip], code[ip + 1] | ext
        ip += 2
        ext = 0
        nm = NAME.get(op)
        if nm is None:
            raise NotImplementedError(f"opcode {op} at {ip - 2} in {co.qualname}")
        if nm in ("CACHE", "NOP", "RESUME"):
            continue
        if nm == "EXTENDED_ARG":
            ext = arg << 8
            continue
        if nm == "POP_TOP":
            stack.pop()
        elif nm == "PUSH_NULL":
            stack.append(_NULL)
        elif nm == "LOAD_CONST":
            stack.append(consts[arg])
        elif nm == "LOAD_FAST":
            v = fast[arg]
            if v is _NULL:
                raise UnboundLocalError(lpn[arg])
            stack.append(v)
        elif nm == "STORE_FAST":
            fast[arg] = stack.pop()
        elif nm == "DELETE_FAST":
            fast[arg] = _NULL
        elif nm == "LOAD_GLOBAL":
            if arg & 1:
                stack.append(_NULL)
            stack.append(lookup_global(names[arg >> 1]))
        elif nm == "STORE_GLOBAL":
            globs[names[arg]] = stack.pop()
        elif nm == "LOAD_NAME":
            n = names[arg]
            stack.append(class_ns[n] if (class_ns is not None and n in class_ns) else lookup_global(n))
        elif nm == "STORE_NAME":
            (class_ns if class_ns is not None else globs)[names[arg]] = stack.pop()
        elif nm == "LOAD_ATTR":
            stack.append(getattr(stack.pop(), names[arg]))
        elif nm == "STORE_ATTR":
            obj = stack.pop()
            setattr(obj, names[arg], stack.pop())
        elif nm == "LOAD_METHOD":
            obj = stack.pop()
            stack.append(_NULL)
            stack.append(getattr(obj, names[arg]))
        elif nm == "PRECALL":
            pass
        elif nm == "KW_NAMES":
            kwnames = consts[arg]
        elif nm == "CALL":
            nk = len(kwnames)
            vals = [stack.pop() for _ in range(arg)][::-1]
            pos, kw = vals[: arg - nk], dict(zip(kwnames, vals[arg - nk:]))
            kwnames = ()
            b, a = stack.pop(), stack.pop()
            if a is _NULL:
                fn = b
            else:          # (callable, self) pair
                fn, pos = a, [b] + pos
            if fn is builtins.super and not pos:
                cls_cell = next(fast[i] for i, n in enumerate(lpn) if n == "__class__")
                stack.append(super(cls_cell.v, first_arg))
            elif fn is builtins.__build_class__:
                stack.append(_build_class(*pos, **kw))
            else:
                stack.append(fn(*pos, **kw))
        elif nm == "CALL_FUNCTION_EX":
            kw = stack.pop() if arg & 1 else {}
            a = stack.pop()
            fn = stack.pop()
            if stack and stack[-1] is _NULL:
                stack.pop()
            stack.append(fn(*a, **kw))
        elif nm == "RETURN_VALUE":
            return stack.pop()
        elif nm == "BINARY_OP":
            r = stack.pop()
            l = stack.pop()
            stack.append(BINARY[arg](l, r))
        elif nm == "BINARY_SUBSCR":
            k = stack.pop()
            stack.append(stack.pop()[k])
        elif nm == "STORE_SUBSCR":
            k = stack.pop()
            obj = stack.pop()
            obj[k] = stack.pop()
        elif nm == "DELETE_SUBSCR":
            k = stack.pop()
            del stack.pop()[k]
        elif nm == "COMPARE_OP":
            r = stack.pop()
            l = stack.pop()
            stack.append(COMPARE[arg](l, r))
        elif nm == "IS_OP":
            r = stack.pop()
            l = stack.pop()
            stack.append((l is not r) if arg else (l is r))
        elif nm == "CONTAINS_OP":
            r = stack.pop()
            l = stack.pop()
            stack.append((l not in r) if arg else (l in r))
        elif nm == "UNARY_NEGATIVE":
            stack.append(-stack.pop())
        elif nm == "UNARY_POSITIVE":
            stack.append(+stack.pop())
        elif nm == "UNARY_NOT":
            stack.append(not stack.pop())
        elif nm == "UNARY_INVERT":
            stack.append(~stack.pop())
        elif nm == "GET_LEN":
            stack.append(len(stack[-1]))
        elif nm == "BUILD_TUPLE":
            vals = stack[len(stack) - arg:] if arg else []
            del stack[len(stack) - arg:]
            stack.append(tuple(vals))
        elif nm == "BUILD_LIST":
            vals = stack[len(stack) - arg:] if arg else []
            del stack[len(stack) - arg:]
            stack.append(list(vals))
        elif nm == "BUILD_SET":
            vals = stack[len(stack) - arg:] if arg else []
            del stack[len(stack) - arg:]
            stack.append(set(vals))
        elif nm == "BUILD_MAP":
            vals = stack[len(stack) - 2 * arg:] if arg else []
            del stack[len(stack) - 2 * arg:]
            stack.append({vals[2 * i]: vals[2 * i + 1] for i in range(arg)})
        elif nm == "BUILD_CONST_KEY_MAP":
            keys = stack.pop()
            vals = stack[len(stack) - arg:]
            del stack[len(stack) - arg:]
            stack.append(dict(zip(keys, vals)))
        elif nm == "BUILD_SLICE":
            step = stack.pop() if arg == 3 else None
            stop = stack.pop()
            start = stack.pop()
            stack.append(slice(start, stop, step))
        elif nm == "BUILD_STRING":
            vals = stack[len(stack) - arg:]
            del stack[len(stack) - arg:]
            stack.append("".join(vals))
        elif nm == "FORMAT_VALUE":
            spec = stack.pop() if (arg & 0x04) else ""
            v = stack.pop()
            conv = arg & 0x03
            v = str(v) if conv == 1 else repr(v) if conv == 2 else ascii(v) if conv == 3 else v
            stack.append(format(v, spec))
        elif nm == "LIST_TO_TUPLE":
            stack.append(tuple(stack.pop()))
        elif nm == "LIST_APPEND":
            v = stack.pop()
            stack[-arg].append(v)
        elif nm == "SET_ADD":
            v = stack.pop()
            stack[-arg].add(v)
        elif nm == "MAP_ADD":
            v = stack.pop()
            k = stack.pop()
            stack[-arg][k] = v
        elif nm == "LIST_EXTEND":
            v = stack.pop()
            stack[-arg].extend(v)
        elif nm == "SET_UPDATE":
            v = stack.pop()
            stack[-arg].update(v)
        elif nm in ("DICT_UPDATE", "DICT_MERGE"):
            v = stack.pop()
            stack[-arg].update(v)
        elif nm == "UNPACK_SEQUENCE":
            vals = list(stack.pop())
            if len(vals) != arg:
                raise ValueError(f"expected {arg} values to unpack, got {len(vals)}")
            stack.extend(vals[::-1])
        elif nm == "COPY":
            stack.append(stack[-arg])
        elif nm == "SWAP":
            stack[-1], stack[-arg] = stack[-arg], stack[-1]
        elif nm == "GET_ITER":
            stack.append(iter(stack.pop()))
        elif nm == "FOR_ITER":
            try:
                stack.append(next(stack[-1]))
            except StopIteration:
                stack.pop()
                ip += 2 * arg
        elif nm == "JUMP_FORWARD":
            ip += 2 * arg
        elif nm == "JUMP_BACKWARD":
            ip -= 2 * arg
        elif nm == "POP_JUMP_FORWARD_IF_FALSE":
            if not stack.pop():
                ip += 2 * arg
        elif nm == "POP_JUMP_FORWARD_IF_TRUE":
            if stack.pop():
                ip += 2 * arg
        elif nm == "POP_JUMP_BACKWARD_IF_FALSE":
            if not stack.pop():
                ip -= 2 * arg
        elif nm == "POP_JUMP_BACKWARD_IF_TRUE":
            if stack.pop():
                ip -= 2 * arg
        elif nm == "POP_JUMP_FORWARD_IF_NONE":
            if stack.pop() is None:
                ip += 2 * arg
        elif nm == "POP_JUMP_FORWARD_IF_NOT_NONE":
            if stack.pop() is not None:
                ip += 2 * arg
        elif nm == "POP_JUMP_BACKWARD_IF_NONE":
            if stack.pop() is None:
                ip -= 2 * arg
        elif nm == "POP_JUMP_BACKWARD_IF_NOT_NONE":
            if stack.pop() is not None:
                ip -= 2 * arg
        elif nm == "JUMP_IF_FALSE_OR_POP":
            if not stack[-1]:
                ip += 2 * arg
            else:
                stack.pop()
        elif nm == "JUMP_IF_TRUE_OR_POP":
            if stack[-1]:
                ip += 2 * arg
            else:
                stack.pop()
        elif nm == "RAISE_VARARGS":
            if arg == 1:
                exc = stack.pop()
                raise exc() if isinstance(exc, type) else exc
            if arg == 2:
                cause = stack.pop()
                exc = stack.pop()
                raise (exc() if isinstance(exc, type) else exc) from cause
            raise RuntimeError("bare raise outside an except block")
        elif nm == "LOAD_ASSERTION_ERROR":
            stack.append(AssertionError)
        elif nm == "LOAD_BUILD_CLASS":
            stack.append(builtins.__build_class__)
        elif nm == "MAKE_FUNCTION":
            fcode = stack.pop()
            clo = stack.pop() if arg & 0x08 else ()
            if arg & 0x04:
                stack.pop()                      # annotations
            kwd = stack.pop() if arg & 0x02 else None
            dfl = stack.pop() if arg & 0x01 else ()
            stack.append(Function(fcode, globs, dfl, kwd, clo))
        elif nm == "MAKE_CELL":
            if not isinstance(fast[arg], Cell):          # (the class body's __class__ cell is placed by _build_class)
                fast[arg] = Cell(fast[arg])
        elif nm == "LOAD_CLOSURE":
            stack.append(fast[arg])
        elif nm in ("LOAD_DEREF", "LOAD_CLASSDEREF"):
            c = fast[arg]
            if c.v is _NULL:
                raise NameError(lpn[arg])
            stack.append(c.v)
        elif nm == "STORE_DEREF":
            fast[arg].v = stack.pop()
        elif nm == "COPY_FREE_VARS":
            n = len(lpn)
            for k in range(arg):
                fast[n - arg + k] = closure[k]
        elif nm == "IMPORT_NAME":
            fromlist = stack.pop()
            level = stack.pop()
            n = names[arg]
            mod = globs["__import__"](n, fromlist, level) if "__import__" in globs else importlib.import_module(n)
            if not fromlist and "__import__" not in globs:
                mod = importlib.import_module(n.split(".")[0])
            stack.append(mod)
        elif nm == "IMPORT_FROM":
            stack.append(getattr(stack[-1], names[arg]))
        else:
            raise NotImplementedError(f"{nm} in {co.qualname}")


def exec_module(path: str, module_name: str, importer=None) -> dict:
    """Run the module-level code of a 3.11 .pyc; returns its globals.  `importer(name, fromlist, level)` resolves
    imports (default: importlib)."""
    co = load_pyc(path)
    globs: Dict[str, Any] = {"__name__": module_name, "__builtins__": builtins}
    if importer is not None:
        globs["__import__"] = importer
    run(co, globs, None)
    globs.pop("__import__", None)
    return globs
