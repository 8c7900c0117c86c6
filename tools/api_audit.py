import ast, os, sys, importlib, json
sys.path.insert(0, "/root/repo")
import warnings; warnings.filterwarnings("ignore")
roots = [("/root/reference/legacy/vescale", "vescale"), ("/root/reference/vescale", "vescale")]
missing_mod, missing_sym = [], []
total_mod = total_sym = 0
for root, pkg in roots:
    for dp, dn, fn in os.walk(root):
        for f in fn:
            if not f.endswith(".py"): continue
            p = os.path.join(dp, f)
            rel = os.path.relpath(p, root)[:-3].replace("/", ".")
            if rel.endswith("__init__"): rel = rel[:-9].rstrip(".")
            mod = pkg + ("." + rel if rel else "")
            if "_pb2" in mod: continue
            try: tree = ast.parse(open(p).read())
            except Exception: continue
            names = []
            for n in tree.body:
                if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and not n.name.startswith("_"): names.append(n.name)
            total_mod += 1
            try:
                m = importlib.import_module(mod)
            except Exception as e:
                missing_mod.append((mod, type(e).__name__, str(e)[:80], len(names))); continue
            for nm in names:
                total_sym += 1
                if not hasattr(m, nm): missing_sym.append(f"{mod}:{nm}")
print("modules", total_mod, "missing", len(missing_mod)); 
for x in sorted(missing_mod): print("  MOD", x)
print("symbols", total_sym, "missing", len(missing_sym))
for x in sorted(missing_sym): print("  SYM", x)
