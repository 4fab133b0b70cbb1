#!/usr/bin/env python3
"""Golden alpha-beta vectors from the reference's own JavaScript: squava.html, the browser
transliteration of the alpha-beta program (evaluator = squavathr's deltaValue with the no2 /
noMiddle2 terms, search = evaluate at entry + cumulative value, i.e. dialect AB-2's alphaBeta).

No JavaScript engine exists in this image, so the page's <script> is translated MECHANICALLY to
Python in memory (statement by statement: var/for/if/switch/function/return, braces -> blocks) and
executed; nothing of the reference is copied into this repo.  Only DOM-free code is kept: the
globals and tables, deltavalue() (squava.html:201-285) and alphabeta() (:333-382).  The root loop of
mymove() (:83-106) touches the DOM only after the search, so it is re-driven here line for line.

Output: tests/golden/ref_squava_html_vectors.json -- per position and max_depth, for every empty
cell: the ply-0 deltavalue ([stop, value]) and the root value mymove() would compare.

TEST INFRASTRUCTURE; run once in the build container.
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_squava_html_vectors.json")

SKIP_FUNCS = {"clork", "mymove", "bookDefender", "bookOffender", "color_winner", "check_winner",
              "resetgame", "cfirst"}


def extract_script(html):
    m = re.search(r"<script[^>]*>(.*?)</script>", html, flags=re.S)
    return m.group(1)


def drop_functions(src):
    """remove the DOM-bound functions wholesale (brace matching)"""
    out, i = [], 0
    for m in re.finditer(r"^function\s+(\w+)\s*\(", src, flags=re.M):
        if m.start() < i:
            continue
        if m.group(1) in SKIP_FUNCS:
            out.append(src[i:m.start()])
            j = src.index("{", m.end())
            depth = 0
            while True:
                if src[j] == "{":
                    depth += 1
                elif src[j] == "}":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            i = j + 1
    out.append(src[i:])
    return "".join(out)


def expr(e):
    e = e.replace("&&", " and ").replace("||", " or ")
    e = re.sub(r"!(?!=)", " not ", e)
    e = re.sub(r"\btrue\b", "True", e)
    e = re.sub(r"\bfalse\b", "False", e)
    e = re.sub(r"new Array\((\d+)\)", r"JSArray([None] * \1)", e)
    e = e.replace("new Array()", "JSArray()")
    e = e.replace(".push(", ".append(")
    return e.strip()


def translate(src):
    # join statements that continue on the next line (an `if (` whose `{` comes later)
    lines, buf = [], ""
    for raw in src.splitlines():
        line = re.sub(r"//.*$", "", raw).rstrip()
        if not line.strip():
            continue
        if buf:
            buf += " " + line.strip()
        else:
            buf = line
        s = buf.strip()
        if s.startswith("if (") and not s.endswith("{"):
            continue
        lines.append(buf)
        buf = ""
    py, stack, bracket = [], [], 0        # stack of open blocks: 'def' 'for' 'if' 'switch:<var>' 'case'
    for line in lines:
        s = line.strip()
        if bracket > 0:                         # inside a multi-line array literal: copy through
            bracket += s.count("[") - s.count("]")
            py.append("    " * len(stack) + "    " + s.rstrip(";"))
            continue
        if s.startswith("}"):
            stack.pop()
            s = s[1:].strip()
            if not s:
                continue
        ind = "    " * len(stack)
        opens = s.endswith("{")
        if opens:
            s = s[:-1].strip()
        kind = "if"
        m = re.match(r"function\s+(\w+)\s*\((.*?)\)$", s)
        if m:
            py.append(ind + "def %s(%s):" % (m.group(1), m.group(2)))
            kind = "def"
        elif re.match(r"for\s*\(var (\w+) = (\w+); \1 < (\w+); \+\+\1\)$", s):
            m = re.match(r"for\s*\(var (\w+) = (\w+); \1 < (\w+); \+\+\1\)$", s)
            py.append(ind + "for %s in range(%s, %s):" % m.groups())
            kind = "for"
        elif re.match(r"for\s*\(var (\w+) in (\w+)\)$", s):
            m = re.match(r"for\s*\(var (\w+) in (\w+)\)$", s)
            py.append(ind + "for %s in range(len(%s)):" % m.groups())
            kind = "for"
        elif re.match(r"else if\s*\((.*)\)$", s):
            py.append(ind + "elif %s:" % expr(re.match(r"else if\s*\((.*)\)$", s).group(1)))
        elif s == "else":
            py.append(ind + "else:")
        elif re.match(r"if\s*\((.*)\)$", s):
            py.append(ind + "if %s:" % expr(re.match(r"if\s*\((.*)\)$", s).group(1)))
        elif re.match(r"switch\s*\((\w+)\)$", s):
            py.append(ind + "if True:")
            kind = "switch:" + re.match(r"switch\s*\((\w+)\)$", s).group(1)
        elif re.match(r"case\s+(-?\w+):$", s):
            var = [k for k in stack if k.startswith("switch:")][-1].split(":")[1]
            py.append(ind + "if %s == %s:" % (var, re.match(r"case\s+(-?\w+):$", s).group(1)))
            stack.append("case")                # its statements sit one level deeper until `break;`
            continue
        elif s == "break;":
            inner = [k for k in stack if k in ("for", "case")][-1]
            if inner == "case" and stack[-1] == "case":
                stack.pop()                     # end of the case
            else:
                py.append(ind + "break")        # a real loop break
            continue
        else:
            s = s.rstrip(";")
            m = re.match(r"var (\w+)$", s)
            if m:
                py.append(ind + "%s = None" % m.group(1))
            elif re.match(r"var (\w+), (\w+)$", s):
                py.append(ind + "%s = %s = None" % re.match(r"var (\w+), (\w+)$", s).groups())
            elif re.match(r"\+\+(\w+)$", s):
                py.append(ind + "%s += 1" % re.match(r"\+\+(\w+)$", s).group(1))
            else:
                s = re.sub(r"^var ", "", s)
                py.append(ind + expr(s))
                bracket = s.count("[") - s.count("]")
        if opens:
            stack.append(kind)
    return "\n".join(py)


class JSArray(list):
    """a JavaScript array: assigning past the end grows it"""
    def __setitem__(self, i, v):
        while isinstance(i, int) and i >= len(self):
            self.append(None)
        list.__setitem__(self, i, v)


def load_reference():
    html = open(os.path.join(REF, "squava.html")).read()
    py = translate(drop_functions(extract_script(html)))
    ns = {"JSArray": JSArray}
    exec(compile(py, "squava.html <script>, translated in memory", "exec"), ns)
    return ns, py


def root_values(ns, board, max_depth):
    """mymove(), squava.html:83-106, without the DOM"""
    ns["max_depth"] = max_depth
    b = ns["board"]
    for i in range(5):
        for j in range(5):
            b[i][j] = board[5 * i + j]
    out = {}
    for j in range(5):
        for i in range(5):
            if b[i][j] == ns["UNSET"]:
                b[i][j] = ns["MAXIMIZER"]
                decisions = ns["deltavalue"](0, i, j, 0)
                if decisions[0]:
                    val = decisions[1]
                else:
                    val = ns["alphabeta"](1, ns["MINIMIZER"], 2 * ns["LOSS"], 2 * ns["WIN"], i, j, decisions[1])
                b[i][j] = ns["UNSET"]
                out[5 * i + j] = [bool(decisions[0]), int(decisions[1]), int(val)]
    return out


def main():
    from squava_b200 import positions
    ns, py = load_reference()
    if "--show" in sys.argv:
        print(py)
        return
    g = np.random.Generator(np.random.PCG64(1020))
    cases = []
    plan = [(16, 6, 12), (14, 6, 10), (12, 6, 6), (12, 4, 10), (10, 4, 10), (18, 8, 8), (20, 8, 6), (8, 4, 4),
            (8, 2, 6), (22, 6, 4), (23, 8, 3), (6, 2, 4), (4, 2, 2), (0, 2, 1), (15, 6, 8), (17, 8, 8),
            (10, 6, 3), (13, 8, 3), (8, 5, 2), (11, 7, 2)]
    for marks, depth, count in plan:
        for _ in range(count):
            x, o = positions.random_position(g, marks)
            if marks % 2:                         # the page's engine is MAXIMIZER to move: make it so
                x, o = o, x
            board = [1 if (x >> c) & 1 else -1 if (o >> c) & 1 else 0 for c in range(25)]
            vals = root_values(ns, board, depth)
            cases.append({"x": x, "o": o, "max_depth": depth, "cells": {str(k): v for k, v in vals.items()}})
            print(marks, depth, len(vals), flush=True)
    scores = [v for row in ns["scores"] for v in row]
    with open(OUT, "w") as f:
        json.dump({"source": "squava.html of the reference, <script> translated in memory and executed by "
                             "oracle/tools/make_ref_golden_js.py", "scores": scores, "cases": cases}, f,
                  separators=(",", ":"))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
