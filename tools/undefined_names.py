"""Poor man's pyflakes (no linter is installed in the image): reports names that are read in a function but bound nowhere in it, in its
enclosing functions, at module level or in builtins.   python tools/undefined_names.py bench.py ant-multi-modal-framework_b200/*.py ..."""
import ast
import builtins
import sys


def bound_names(node):
    out = set()
    for n in ast.walk(node):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
        elif isinstance(n, ast.arg):
            out.add(n.arg)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                out.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            out.update(n.names)
    return out


def check(path):
    tree = ast.parse(open(path).read(), path)
    mod = bound_names(tree) | set(dir(builtins)) | {"__file__", "__name__"}
    bad = []

    def visit(fn, outer):
        local = bound_names(fn) | outer
        for n in ast.walk(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in local:
                bad.append((path, n.lineno, n.id))

    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef)):
            visit(n, mod)
    return bad


if __name__ == "__main__":
    problems = [p for f in sys.argv[1:] for p in check(f)]
    for p in sorted(set(problems)):
        print("%s:%d: undefined name %r" % p)
    sys.exit(1 if problems else 0)
