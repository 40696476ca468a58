"""Line-overlap of the product's Python sources with the reference tree (copy hygiene check).

Docstrings, comments and blank lines are stripped, whitespace is normalised; a product line counts as
shared when the identical normalised line occurs anywhere in /root/reference/pyqmc.  Lines shorter than
MINLEN characters (``return x``, ``else:`` ...) are ignored on both sides."""
import ast
import io
import os
import sys
import tokenize

MINLEN = 12


def code_lines(path):
    src = open(path).read()
    drop = set()
    try:
        tree = ast.parse(src)
        for node in ast.walk(tree):
            if isinstance(node, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef, ast.Module)):
                b = node.body
                if b and isinstance(b[0], ast.Expr) and isinstance(getattr(b[0], "value", None), ast.Constant) \
                        and isinstance(b[0].value.value, str):
                    drop.update(range(b[0].lineno, b[0].end_lineno + 1))
    except SyntaxError:
        pass
    comments = {}
    try:
        for tok in tokenize.generate_tokens(io.StringIO(src).readline):
            if tok.type == tokenize.COMMENT:
                comments[tok.start[0]] = tok.start[1]
    except tokenize.TokenError:
        pass
    out = []
    for i, line in enumerate(src.splitlines(), 1):
        if i in drop:
            continue
        if i in comments:
            line = line[: comments[i]]
        norm = "".join(line.split())
        if len(norm) >= MINLEN:
            out.append(norm)
    return out


def reference_lines(root):
    ref = set()
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                ref.update(code_lines(os.path.join(d, f)))
    return ref


def report(product_dir, ref_root="/root/reference/pyqmc"):
    ref = reference_lines(ref_root)
    rows = []
    for d, _, files in os.walk(product_dir):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                lines = code_lines(p)
                if lines:
                    shared = sum(1 for l in lines if l in ref)
                    rows.append((shared / len(lines), shared, len(lines), os.path.relpath(p, product_dir)))
    return sorted(rows, reverse=True)


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for frac, shared, n, name in report(os.path.join(here, sys.argv[1] if len(sys.argv) > 1 else "pyqmc_b200")):
        print(f"{100 * frac:5.1f}%  {shared:4d}/{n:4d}  {name}")
