"""Crude undefined-name check (no pyflakes in the image): every Name loaded inside a function must be
a parameter or a name bound in that function or an enclosing one, a module-level name or a builtin.
Usage: python tools/check_names.py file.py ..."""
import ast
import builtins
import sys

SCOPES = (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)


def _bound_here(fn):
    """Names bound directly in the scope of fn (not descending into nested scopes, except their names)."""
    names = set()
    a = fn.args
    for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
        names.add(arg.arg)
    stack = list(ast.iter_child_nodes(fn))
    while stack:
        n = stack.pop()
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(n.name)
            continue
        if isinstance(n, ast.Lambda):
            continue
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            names.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for al in n.names:
                names.add((al.asname or al.name).split(".")[0])
        elif isinstance(n, ast.ExceptHandler) and n.name:
            names.add(n.name)
        stack.extend(ast.iter_child_nodes(n))
    return names


def check(path):
    tree = ast.parse(open(path).read(), path)
    mod = set(dir(builtins)) | {"__file__", "__name__"}
    stack = list(tree.body)
    while stack:                                   # module-level bindings (also inside if/try/for/with blocks)
        n = stack.pop()
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            mod.add(n.name)
            continue
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for al in n.names:
                mod.add((al.asname or al.name).split(".")[0])
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
            mod.add(n.id)
        stack.extend(ast.iter_child_nodes(n))
    bad = []

    def visit(node, scope):
        for child in ast.iter_child_nodes(node):
            if isinstance(child, SCOPES):
                visit(child, scope | _bound_here(child))
            elif isinstance(child, ast.ClassDef):
                visit(child, scope)                # class bodies: methods see the enclosing function scope only
            else:
                if isinstance(child, ast.Name) and isinstance(child.ctx, ast.Load) and scope is not None \
                        and child.id not in scope and child.id not in mod:
                    bad.append((path, child.lineno, child.id))
                visit(child, scope)

    for node in tree.body:
        if isinstance(node, SCOPES):
            visit(node, _bound_here(node))
        elif isinstance(node, ast.ClassDef):
            cls_names = {n.name for n in node.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
            for n in ast.walk(node):
                if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
                    cls_names.add(n.id)
            for sub in node.body:
                if isinstance(sub, SCOPES):
                    visit(sub, _bound_here(sub))
    return bad


if __name__ == "__main__":
    out = []
    for p in sys.argv[1:]:
        out += check(p)
    for b in sorted(set(out)):
        print("%s:%d: undefined name %s" % b)
    sys.exit(1 if out else 0)
