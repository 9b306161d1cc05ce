"""Prints what the BFS-bistride spec (oracle/bistride_oracle.py) was read from: the code-object
tree, names, constants and local-variable load order of the two bytecode-only modules of the reference.
TEST INFRASTRUCTURE / documentation only; needs /root/reference (not present on the GPU box).  marshal of CPython 3.11
code objects loads under 3.12, the bytecode itself is not executed.

    python oracle/decode_bistride_pyc.py [/root/reference]
"""
import marshal
import sys
import types

LOAD_FAST, STORE_FAST, BINARY_OP = 124, 125, 122      # CPython 3.11 opcode numbers
BINOPS = {0: "+", 5: "*", 6: "%", 10: "-", 1: "&", 13: "+="}


def walk(co, depth=0):
    pad = "  " * depth
    print(f"{pad}{co.co_name} (orig :{co.co_firstlineno}) args={co.co_varnames[:co.co_argcount]}")
    print(f"{pad}  names : {co.co_names}")
    print(f"{pad}  consts: {[c for c in co.co_consts if not isinstance(c, (types.CodeType, str)) or (isinstance(c, str) and len(c) < 40)]}")
    ops = []
    code = co.co_code
    for i in range(0, len(code), 2):
        op, arg = code[i], code[i + 1]
        if op in (LOAD_FAST, STORE_FAST) and arg < len(co.co_varnames):
            ops.append(("<-" if op == STORE_FAST else "") + co.co_varnames[arg])
        elif op == BINARY_OP and arg in BINOPS:
            ops.append(BINOPS[arg])
    print(f"{pad}  locals/binary ops in order: {' '.join(ops)}")
    for c in co.co_consts:
        if isinstance(c, types.CodeType):
            walk(c, depth + 1)


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    for name in ("bistride_ops", "bsms_mgn"):
        path = f"{root}/models/__pycache__/{name}.cpython-311.pyc"
        print("=" * 20, path)
        walk(marshal.loads(open(path, "rb").read()[16:]))
