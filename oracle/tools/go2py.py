#!/usr/bin/env python3
"""A mechanical translator for the small subset of Go the reference's search code is written in.

TEST INFRASTRUCTURE.  There is no Go toolchain in this image; to obtain OUTPUT of the reference's
own Go code, selected top-level declarations (tables, functions, methods) are translated statement
by statement to Python IN MEMORY and executed (oracle/tools/make_ref_golden_go.py).  Nothing of
the reference is copied into this repo.

Subset handled: const/var tables (composite literals of ints), `x := e`, `a, b := f()`, `=`, `+=`,
`++`, `if / else if / else`, `for i := a; i < b; i++`, `for i, v := range e`, `for _, v := range e`,
`for cond {`, `switch v { case X: ... }` (no fallthrough), `break`/`continue`, `return` with named
results, methods with pointer receivers, `append(s, v)`, `!`, `&&`, `||`, `nil`, `true`, `false`.
Integer division: every `/` in the translated code acts on ints whose quotient is exact (sum/3 with
sum = +-3), and is emitted as int(a / b), i.e. truncation toward zero like Go.
"""
import re


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def top_level_blocks(src):
    """-> list of (kind, name, text) for func / var / const / type declarations"""
    out, i, n = [], 0, len(src)
    pat = re.compile(r"^(func|var|const|type)\b", re.M)
    while True:
        m = pat.search(src, i)
        if not m:
            break
        j = m.end()
        # find the end: either end of line (no brace/paren opened) or the matching close
        depth, k, opened = 0, j, False
        while k < n:
            c = src[k]
            if c in "{(":
                depth += 1
                opened = True
            elif c in "})":
                depth -= 1
            elif c == "\n" and depth == 0:
                if opened or not src[j:k].rstrip().endswith(("=", ",")):
                    break
            k += 1
        text = src[m.start():k]
        kind = m.group(1)
        name = None
        if kind == "func":
            mm = re.match(r"func\s+(?:\((\w+)\s+\*?(\w+)\)\s*)?(\w+)", text)
            name = (mm.group(2) + "_" if mm.group(2) else "") + mm.group(3)
        elif kind in ("var", "type"):
            mm = re.match(r"(?:var|type)\s+(\w+)", text)
            name = mm.group(1) if mm else None
        out.append((kind, name, text))
        i = k
    return out


def names_of(params):
    """'ply int, x, y int, bd *Board' -> ['ply', 'x', 'y', 'bd']"""
    names = []
    for piece in [p.strip() for p in params.split(",") if p.strip()]:
        names.append(piece.split()[0])
    return names


def expr(e):
    e = e.strip()
    e = e.replace("&&", " and ").replace("||", " or ")
    e = re.sub(r"!(?!=)", " not ", e)
    e = re.sub(r"\btrue\b", "True", e)
    e = re.sub(r"\bfalse\b", "False", e)
    e = re.sub(r"\bnil\b", "None", e)
    e = re.sub(r"\[\]\w+\{\}", "[]", e)                    # []int{}
    e = re.sub(r"&\(([^()]*)\)", r"\1", e)             # &(curr.bd) -> curr.bd
    e = re.sub(r"&(\w)", r"\1", e)                      # &bd -> bd
    e = re.sub(r"\bappend\(\s*(.+),\s*([^,()]+)\.\.\.\)$", r"list(\1) + list(\2)", e)   # append(a, b...)
    e = re.sub(r"\bappend\(\s*(.+),\s*([^,()]+)\)$", r"list(\1) + [\2]", e)
    # Go integer division truncates toward zero
    e = re.sub(r"(-?\w+(?:\[[^\]]*\])*)\s*/\s*(\w+)", r"int(\1 / \2)", e)
    return e


def literal(text):
    """Go composite literal of ints -> Python list expression"""
    text = re.sub(r"(\[\d*\])+\w+\s*\{", "{", text)   # [][]int{ , [2]int{ , [5][5]int{
    text = text.replace("{", "[").replace("}", "]")
    text = re.sub(r"(-?\d+)\s*-\s*(\d+)", lambda m: str(int(m.group(1)) - int(m.group(2))), text)
    return text


def translate_var(text):
    m = re.match(r"(?:var|const)\s+(\w+)[^=\n]*=\s*(.*)$", text, flags=re.S)
    if m:
        rhs = m.group(2).strip()
        if "{" in rhs:
            return "%s = %s" % (m.group(1), literal(rhs))
        return "%s = %s" % (m.group(1), expr(rhs))
    m = re.match(r"var\s+(\w+)\s+((?:\[\d+\])+)(.*)$", text.strip())
    if m:                                                # var indexed [5][5][][][]int : zero value
        dims = [int(d) for d in re.findall(r"\[(\d+)\]", m.group(2))]
        leaf = "[]" if "[]" in m.group(3) else "0"
        e = leaf
        for d in reversed(dims):
            e = "[%s for _ in range(%d)]" % (e, d)
        return "%s = %s" % (m.group(1), e)
    m = re.match(r"var\s+(\w+)\s+\w+$", text.strip())
    if m:
        return "%s = 0" % m.group(1)
    return "# skipped: " + text.splitlines()[0]


ARRAY_FIELDS = set()     # struct fields of array type seen so far


def translate_type(text):
    """type T struct { a int; b [25][2]int; c bool } -> a class whose instances hold zero values"""
    m = re.match(r"type\s+(\w+)\s+struct\s*\{(.*)\}", text, flags=re.S)
    if not m:
        return "# skipped: " + text.splitlines()[0]
    lines = ["class %s:" % m.group(1), "    def __init__(self):"]
    for fld in m.group(2).splitlines():
        parts = fld.split()
        if len(parts) < 2:
            continue
        typ = parts[-1]
        names = [n.strip(",") for n in parts[:-1]]
        dims = [int(d) for d in re.findall(r"\[(\d+)\]", typ)]
        base = re.sub(r"(\[\d*\])+", "", typ)
        zero = ("[]" if typ.startswith("[]") else "0.0" if base == "float64" else "False" if base == "bool"
                else "None" if base.startswith(("*", "func")) else "0")
        for d in reversed(dims):
            zero = "[%s for _ in range(%d)]" % (zero, d)
        for n in names:
            lines.append("        self.%s = %s" % (n, zero))
            if dims:
                ARRAY_FIELDS.add(n)
    return "\n".join(lines)


def translate_const_block(text):
    out = []
    iota = None
    for pos, line in enumerate(l for l in text.splitlines() if l.strip() and l.strip() not in ("const (", ")")):
        m = re.match(r"\s*(?:const\s+)?(\w+)\s*=\s*(-?\d+)\s*$", line)      # plain integer constants
        if m:
            out.append("%s = %s" % (m.group(1), m.group(2)))
            iota = None
            continue
        m = re.match(r"\s*(\w+)\s*=\s*iota\s*$", line)                       # NAME = iota, then bare names
        if m:
            iota = pos
            out.append("%s = %d" % (m.group(1), pos))
            continue
        m = re.match(r"\s*(\w+)\s*$", line)
        if m and iota is not None:
            out.append("%s = %d" % (m.group(1), pos))
    return "\n".join(out)


def translate_func(text, globals_written=()):
    m = re.match(r"func\s+(?:\((\w+)\s+\*?(\w+)\)\s*)?(\w+)\s*\((.*?)\)\s*(\([^)]*\)|[\w\[\]\*]*)\s*\{", text, flags=re.S)
    recv, rtype, fname, params, results = m.groups()
    name = (rtype + "_" if rtype else "") + fname
    args = ([recv] if recv else []) + names_of(params)
    named = []
    if results.startswith("("):
        inner = results[1:-1]
        pieces = [p.strip() for p in inner.split(",") if p.strip()]
        if pieces and len(pieces[0].split()) > 1 or any(len(p.split()) > 1 for p in pieces):
            named = names_of(inner)
    body = text[m.end():text.rindex("}")]
    py = ["def %s(%s):" % (name, ", ".join(args))]
    if globals_written:
        py.append("    global " + ", ".join(globals_written))
    for r in named:
        py.append("    %s = 0" % r)
    stack = ["def"]
    labels = []          # (label, depth of the labelled loop) for `break LABEL`
    pending_label = None
    lines = []
    buf = ""
    for raw in body.splitlines():                         # join continuation lines (open parens)
        line = raw.rstrip()
        if not line.strip():
            continue
        buf = (buf + " " + line.strip()) if buf else line
        tail = buf.rstrip()
        if buf.count("(") > buf.count(")") or tail.endswith((",", "||", "&&")) or (tail.endswith("+") and not tail.endswith("++")):
            continue
        lines.append(buf)
        buf = ""
    for line in lines:
        s = line.strip()
        while s.startswith("}"):
            if stack[-1] == "case":          # the brace closes the switch: its last case ends too
                stack.pop()
            closed = stack.pop()
            if closed.startswith("forpost:"):
                py.append("    " * (len(stack) + 1) + closed.split(":", 1)[1])
            if closed == "for" or closed.startswith("forpost:"):
                if labels and labels[-1][1] == len(stack):
                    labels.pop()                              # the labelled loop itself ended
                for lab, depth in labels:                     # a loop nested in a labelled one ended
                    if depth < len(stack):
                        py.append("    " * len(stack) + "if _brk_%s: break" % lab)
            s = s[1:].strip()
        if stack and stack[-1] == "case" and (s.startswith("case ") or s.startswith("default:")):
            stack.pop()
        if not s:
            continue
        ind = "    " * len(stack)
        if re.match(r"^[A-Z]\w*:$", s):                    # a label: belongs to the loop that follows
            pending_label = s[:-1]
            continue
        opens = s.endswith("{")
        if opens:
            s = s[:-1].strip()
        if pending_label and s.startswith("for"):
            py.append(ind + "_brk_%s = False" % pending_label)
            labels.append((pending_label, len(stack)))
            pending_label = None
        if re.match(r"^break\s+\w+$", s):
            py.append(ind + "_brk_%s = True" % s.split()[1])
            py.append(ind + "break")
            continue
        kind = "if"
        mm = re.match(r"for\s+(\w+)\s*:=\s*(.+?);\s*\1\s*<\s*(.+?);\s*\1\+\+$", s)
        if mm:
            py.append(ind + "for %s in range(%s, %s):" % (mm.group(1), expr(mm.group(2)), expr(mm.group(3)))); kind = "for"
        elif re.match(r"for\s+(\w+),\s*(\w+)\s*:=\s*range\s+(.+)$", s):
            a, b, c = re.match(r"for\s+(\w+),\s*(\w+)\s*:=\s*range\s+(.+)$", s).groups()
            if a == "_":
                py.append(ind + "for %s in %s:" % (b, expr(c)))
            else:
                py.append(ind + "for %s, %s in enumerate(%s):" % (a, b, expr(c)))
            kind = "for"
        elif re.match(r"for\s+(\w+)\s*:=\s*range\s+(.+)$", s):
            a, c = re.match(r"for\s+(\w+)\s*:=\s*range\s+(.+)$", s).groups()
            py.append(ind + "for %s in range(len(%s)):" % (a, expr(c))); kind = "for"
        elif re.match(r"for\s*(.*?);\s*(.+?);\s*(.+)$", s) and opens:
            init, cond, post = re.match(r"for\s*(.*?);\s*(.+?);\s*(.+)$", s).groups()
            if init.strip():
                py.append(ind + expr(init.replace(":=", "=")))
            py.append(ind + "while %s:" % expr(cond))
            post = post.strip()
            post = (post[:-2] + " += 1") if post.endswith("++") else expr(post)
            kind = "forpost:" + post          # emitted when the block closes (no `continue` inside)
        elif s == "for":
            py.append(ind + "while True:"); kind = "for"
        elif re.match(r"for\s+(.+)$", s) and opens:
            py.append(ind + "while %s:" % expr(re.match(r"for\s+(.+)$", s).group(1))); kind = "for"
        elif re.match(r"else if\s+(.+)$", s):
            py.append(ind + "elif %s:" % expr(re.match(r"else if\s+(.+)$", s).group(1)))
        elif s == "else":
            py.append(ind + "else:")
        elif re.match(r"if\s+(.+)$", s) and opens:
            py.append(ind + "if %s:" % expr(re.match(r"if\s+(.+)$", s).group(1)))
        elif re.match(r"switch\s+(.+)$", s):
            tmp = "_sw%d" % len(py)                       # Go evaluates the tag once
            py.append(ind + "%s = %s" % (tmp, expr(re.match(r"switch\s+(.+)$", s).group(1))))
            py.append(ind + "for _once in (0,):")          # one pass; a `break` in a case leaves the switch
            kind = "switch:" + tmp
        elif re.match(r"case\s+(.+):$", s):
            var = [k for k in stack if k.startswith("switch:")][-1].split(":", 1)[1]
            py.append(ind + "if %s == %s:" % (var, expr(re.match(r"case\s+(.+):$", s).group(1))))
            stack.append("case")
            continue
        elif s in ("break", "continue"):
            inner = [("for" if k.startswith("forpost:") else k) for k in stack if k in ("for", "case") or k.startswith("forpost:")]
            if s == "continue" and inner and inner[-1] == "case":
                raise NotImplementedError("continue inside a switch case")
            py.append(ind + s)                             # in a case: leaves the one-pass loop of the switch
        elif s == "return":
            py.append(ind + "return " + (", ".join(named) if named else ""))
        elif s.startswith("return "):
            py.append(ind + "return " + ", ".join(expr(p) for p in split_top(s[7:])))
        elif re.match(r"(\w+)\+\+$", s):
            py.append(ind + "%s += 1" % s[:-2])
        elif re.match(r"([\w\.\[\]]+)\+\+$", s):
            py.append(ind + "%s += 1" % s[:-2])
        elif re.match(r"var\s+(\w+)\s+\*\w+$", s):
            py.append(ind + "%s = None" % re.match(r"var\s+(\w+)", s).group(1))
        elif re.match(r"var\s+(\w+)\s+\[\]\[\d+\]\w+$", s):              # a nil slice of arrays
            py.append(ind + "%s = []" % re.match(r"var\s+(\w+)", s).group(1))
        elif re.match(r"var\s+(\w+)\s+\[\]\w+$", s):
            py.append(ind + "%s = []" % re.match(r"var\s+(\w+)", s).group(1))
        elif re.match(r"var\s+(\w+(?:\s*,\s*\w+)+)\s+(int|float64)$", s):
            for nm in re.match(r"var\s+(.+?)\s+(?:int|float64)$", s).group(1).split(","):
                py.append(ind + "%s = 0" % nm.strip())
        elif re.match(r"var\s+(\w+)\s+(int|bool|float64|string)$", s):
            py.append(ind + "%s = 0" % re.match(r"var\s+(\w+)", s).group(1))
        elif re.match(r"var\s+(\w+)\s+(\w+)$", s):
            mm = re.match(r"var\s+(\w+)\s+(\w+)$", s)
            py.append(ind + "%s = %s()" % mm.groups())          # a struct: its zero value
        else:
            mm = re.match(r"(.+?)\s*(:=|\+=|-=|=)\s*(.+)$", s)
            if mm and not re.match(r".*[=!<>]$", mm.group(1)):
                lhs, op, rhs = mm.groups()
                op = "=" if op == ":=" else op
                fm = re.match(r"^\w+\.(\w+)$", rhs.strip())
                if op == "=" and fm and fm.group(1) in ARRAY_FIELDS:     # Go arrays are values: assignment copies
                    rhs = "list(%s)" % rhs.strip()
                lhs = ", ".join(x.strip() for x in lhs.split(","))
                py.append(ind + "%s %s %s" % (lhs, op, expr(rhs)))
            else:
                py.append(ind + expr(s))
        if opens:
            stack.append(kind)
    return "\n".join(py)


def split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def translate(src, want, globals_written=None):
    """Translate the named top-level declarations of one Go file.  want: iterable of names
    (methods as Type_name).  -> python source"""
    globals_written = globals_written or {}
    src = strip_comments(src)
    out = []
    for kind, name, text in top_level_blocks(src):
        if kind == "const" and re.match(r"const\s*\(", text):
            out.append(translate_const_block(text))
        elif kind == "const":
            out.append(translate_const_block(text))
        elif name in want:
            if kind == "var":
                out.append(translate_var(text))
            elif kind == "type":
                out.append(translate_type(text))
            elif kind == "func":
                out.append(translate_func(text, globals_written.get(name, ())))
    return "\n\n".join(out)


def attach_methods(ns):
    """make `obj.Method(args)` work: every translated Type_method becomes a method of class Type"""
    for name, fn in list(ns.items()):
        if "_" in name and callable(fn):
            t, m = name.split("_", 1)
            if isinstance(ns.get(t), type):
                setattr(ns[t], m, fn)
