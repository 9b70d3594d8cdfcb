"""Minimal lowering of md_script property statements to libmdgpu property descriptors.

In VIAMD the script front-end (tokenizer, parser, static type check, static evaluation of selections) is mdlib's own
md_script.c and stays unchanged; INTEGRATION.md shows the shim that walks a compiled md_script_ir_t and emits the same
descriptors. This module exists so that the tests and bench.py can be written the way the reference's own tests are
(mdlib/unittest/test_script.c:1259-1273: `prop1 = rdf(element('C'), element('O'), 20.0);`) without the reference on the box.

Supported statements:  ident = proc(args);   with proc in
    rdf(sel, sel, cutoff | min:max)   sdf(residue(a:b) | sel-array, sel, cutoff)   density_x|_y|_z(sel)
    distance(i, j)   angle(i, j, k)   dihedral(i, j, k, l)            (1-based atom indices as in md_script)
Selections (evaluated once, statically, to ascending atom index lists — md_script.c:5492-5524):
    all | element('O') | name('C2*') | resname('SOL') | atom(a:b) | residue(a:b) | `and` / `or` / `not` of these
`residue(a:b)` used as the first argument of sdf() yields one structure per residue (bitfield array semantics).
"""
from __future__ import annotations

import fnmatch
import re
from typing import List

import numpy as np

from . import api

_TOK = re.compile(r"\s*(?:(\d+\.\d*|\.\d+|\d+)|([A-Za-z_][A-Za-z_0-9]*)|'([^']*)'|\"([^\"]*)\"|(.))")


class ScriptError(ValueError):
    pass


def _tokens(src: str):
    out = []
    for m in _TOK.finditer(src):
        num, ident, s1, s2, ch = m.groups()
        if num is not None: out.append(("num", num))
        elif ident is not None: out.append(("id", ident))
        elif s1 is not None: out.append(("str", s1))
        elif s2 is not None: out.append(("str", s2))
        elif ch and not ch.isspace(): out.append(("ch", ch))
    return out


class _Parser:
    def __init__(self, toks, system: api.System):
        self.t, self.i, self.sys = toks, 0, system

    def peek(self): return self.t[self.i] if self.i < len(self.t) else ("eof", "")
    def next(self): tok = self.peek(); self.i += 1; return tok

    def expect(self, kind, val=None):
        tok = self.next()
        if tok[0] != kind or (val is not None and tok[1] != val):
            raise ScriptError(f"expected {val or kind}, got {tok[1]!r}")
        return tok

    # ---- selections -> boolean masks
    _static_seen = False
    _within = None   # (min, max, selection) of the one dynamic within() met while parsing a selection expression, see dyn_selection()

    def sel_or(self):
        m = self.sel_and()
        while self.peek() == ("id", "or"):
            if self._within is not None: raise ScriptError("within() is lowered as `selection and within(...)` only, not under `or`")
            self.next(); m = m | self.sel_and()
            if self._within is not None: raise ScriptError("within() is lowered as `selection and within(...)` only, not under `or`")
        return m

    def sel_and(self):
        m = self.sel_not()
        while self.peek() == ("id", "and"):
            self.next(); m = m & self.sel_not()
        return m

    def sel_not(self):
        if self.peek() == ("id", "not"):
            had = self._within
            self.next(); m = ~self.sel_not()
            if self._within is not had: raise ScriptError("`not within(...)` is not lowered")
            return m
        return self.sel_atom()

    def dyn_selection(self):
        """`within([min:]max, sel)` or `static and within(...)` (either order) -> (min, max, within's selection, static side's atoms or None);
        the dynamic part is evaluated per frame on the device, the static side becomes its AND mask (_and md_script_functions.inl:1975)"""
        self._within = None; self._static_seen = False
        m = self.sel_or()
        if self._within is None: raise ScriptError("a within(...) expression was expected")
        lo, hi, sel = self._within; self._within = None
        return lo, hi, sel, (np.nonzero(m)[0].astype(np.int32) if self._static_seen else None)

    def _range(self, count):
        """a | a:b | : (1-based inclusive, as in md_script) -> python slice bounds (0-based, exclusive end)"""
        lo, hi = 1, count
        if self.peek()[0] == "num":
            lo = int(float(self.next()[1])); hi = lo
        if self.peek() == ("ch", ":"):
            self.next(); hi = count
            if self.peek()[0] == "num": hi = int(float(self.next()[1]))
        return max(lo - 1, 0), min(hi, count)

    def sel_atom(self):
        n = self.sys.num_atoms
        tok = self.next()
        if tok == ("ch", "("):
            m = self.sel_or(); self.expect("ch", ")"); return m
        if tok[0] != "id":
            raise ScriptError(f"unexpected token {tok[1]!r} in selection")
        f = tok[1]
        if f == "within":   # only inside dyn_selection(): stands for "every atom" in the static mask, the device supplies the real set
            if self._within is not None: raise ScriptError("one within() per expression")
            self.expect("ch", "("); lo, hi = self.radius(); self.expect("ch", ",")
            seen = self._static_seen; sel = self.selection(); self._static_seen = seen; self.expect("ch", ")")   # within's own argument is not the static side; FLAG_FLATTEN (:673): an array of selections is their union
            self._within = (lo, hi, sel); return np.ones(n, bool)
        self._static_seen = True
        if f == "all": return np.ones(n, bool)
        self.expect("ch", "(")
        if f in ("element", "name", "label", "resname"):
            pats = [self.expect("str")[1]]
            while self.peek() == ("ch", ","):
                self.next(); pats.append(self.expect("str")[1])
            self.expect("ch", ")")
            if f == "element":
                if self.sys.element is None: raise ScriptError("system has no element data")
                src = np.asarray(self.sys.element); return np.isin(np.char.upper(src.astype(str)), [p.upper() for p in pats])
            if f == "resname":
                if self.sys.resname is None: raise ScriptError("system has no residue data")
                rn = np.asarray(self.sys.resname); hit = np.zeros(len(rn), bool)
                for p in pats: hit |= np.array([fnmatch.fnmatchcase(r, p) for r in rn])
                return np.repeat(hit, np.diff(self.sys.res_atom_offset))
            if self.sys.name is None: raise ScriptError("system has no atom names")
            nm = np.asarray(self.sys.name); hit = np.zeros(n, bool)
            for p in pats: hit |= np.array([fnmatch.fnmatchcase(a, p) for a in nm])
            return hit
        if f == "atom":
            lo, hi = self._range(n); self.expect("ch", ")")
            m = np.zeros(n, bool); m[lo:hi] = True; return m
        if f == "residue":
            off = np.asarray(self.sys.res_atom_offset); lo, hi = self._range(len(off) - 1); self.expect("ch", ")")
            m = np.zeros(n, bool); m[off[lo]:off[hi]] = True; return m
        raise ScriptError(f"unsupported selection '{f}'")

    def selection(self) -> np.ndarray:
        return np.nonzero(self.sel_or())[0].astype(np.int32)

    def structures(self) -> np.ndarray:
        """first argument of sdf(): residue(a:b) -> one structure per residue; otherwise a single structure"""
        save = self.i
        if self.peek() == ("id", "residue"):
            self.next(); self.expect("ch", "(")
            off = np.asarray(self.sys.res_atom_offset); lo, hi = self._range(len(off) - 1); self.expect("ch", ")")
            if self.peek() in (("ch", ","),):
                sizes = np.diff(off[lo:hi + 1])
                if len(sizes) == 0 or np.any(sizes != sizes[0]):
                    raise ScriptError("The supplied reference bitfields are not identical")   # _sdf validation :5837
                return np.stack([np.arange(off[r], off[r + 1], dtype=np.int32) for r in range(lo, hi)])
            self.i = save
        s = self.selection()
        return s.reshape(1, -1)

    def groups(self):
        """first argument of rdf(): residue(a:b) covering more than one residue is an ARRAY of bitfields -> one group per residue
        (centre-of-mass references, compute_rdf :5274); anything else is a plain selection (returns None, position unchanged)"""
        save = self.i
        if self.peek() == ("id", "residue"):
            self.next(); self.expect("ch", "(")
            off = np.asarray(self.sys.res_atom_offset); lo, hi = self._range(len(off) - 1); self.expect("ch", ")")
            if self.peek() == ("ch", ",") and hi - lo > 1:
                return [np.arange(off[r], off[r + 1], dtype=np.int32) for r in range(lo, hi)]
        self.i = save
        return None

    def single_selection(self) -> np.ndarray:
        """a selection argument where the reference would treat an ARRAY of selections differently from their union (one position per
        selection: coordinate_extract :1414): residue(a:b) over several residues standing alone is rejected instead of being flattened"""
        save = self.i
        if self.peek() == ("id", "residue"):
            self.next(); self.expect("ch", "(")
            lo, hi = self._range(len(np.asarray(self.sys.res_atom_offset)) - 1); self.expect("ch", ")")
            if self.peek() in (("ch", ","), ("ch", ")")) and hi - lo > 1:
                raise ScriptError("an array of selections as one argument (one centre of mass per selection) is not lowered")
        self.i = save
        return self.selection()

    def _has_within_before_comma(self) -> bool:
        """does the argument that starts here (up to its top-level `,` or `)`) contain a within(...) call"""
        depth = 0
        for t in self.t[self.i:]:
            if t == ("ch", "("): depth += 1
            elif t == ("ch", ")"):
                if depth == 0: return False
                depth -= 1
            elif t == ("ch", ",") and depth == 0: return False
            elif t == ("id", "within"): return True
        return False

    def groups_or_selection(self):
        """residue(a:b) over several residues standing alone -> list of index arrays (array of selections); anything else -> one index array"""
        save = self.i
        if self.peek() == ("id", "residue"):
            self.next(); self.expect("ch", "(")
            off = np.asarray(self.sys.res_atom_offset); lo, hi = self._range(len(off) - 1); self.expect("ch", ")")
            if self.peek() in (("ch", ","), ("ch", ")")) and hi - lo > 1:
                return [np.arange(off[r], off[r + 1], dtype=np.int32) for r in range(lo, hi)]
        self.i = save
        return self.selection()

    def sel_or_within(self, single=False):
        """a selection argument that may be a dynamic one: index array, or api.Within for `within([min:]max, sel)` [and static]"""
        if self._has_within_before_comma():
            lo, hi, sel, cand = self.dyn_selection()
            return api.Within(hi, sel, lo, cand)
        return self.single_selection() if single else self.selection()

    def number(self) -> float:
        return float(self.expect("num")[1])

    def radius(self):
        """radius | min:max -> (min, max)"""
        a = self.number()
        if self.peek() == ("ch", ":"):
            self.next(); return a, self.number()
        return 0.0, a

    def index(self, flatten=False):
        """argument of distance/angle/dihedral/com: a 1-based atom index (-> int), a selection (-> index array, centre of mass) or an ARRAY of
        selections (residue(a:b) over several residues standing alone -> list of index arrays: the centre of the selections' centres,
        coordinate_extract_com :1826-1842). distance() is FLAG_FLATTEN (md_script_functions.inl:680): its arguments, com(...) inside it included,
        are evaluated flattened -> the union."""
        if self.peek()[0] == "num":
            v = int(float(self.expect("num")[1])) - 1   # md_script atom indices are 1-based
            self.arg_meta.append(("int", v)); return v
        if self.peek() == ("id", "com"):   # com(x) as an argument contributes the position x itself would (_com :4726 = coordinate_extract_com)
            self.next(); self.expect("ch", "("); a = self.index(flatten); self.expect("ch", ")")
            return a
        if self._has_within_before_comma(): self.arg_meta.append(("other",)); return self.sel_or_within()
        if self.peek() == ("id", "atom"):   # atom(a:b) standing alone: relative to the context when the expression is evaluated `in` contexts
            save = self.i; self.next(); self.expect("ch", "("); lo, hi = self._range(self.sys.num_atoms); self.expect("ch", ")")
            if self.peek() in (("ch", ","), ("ch", ")")): self.arg_meta.append(("atomrange", lo, hi))
            else: self.arg_meta.append(("other",))
            self.i = save
            return self.selection()
        ctx_relative = self._arg_mentions(("atom", "residue"))   # inside `in` contexts these count from the context's first atom / residue: only the shim
        a = self.selection() if flatten else self.groups_or_selection()   # (which asks mdlib's own evaluator per context) lowers such arguments
        self.arg_meta.append(("other",) if (isinstance(a, list) or ctx_relative) else ("sel", a))
        return a

    def _arg_mentions(self, names) -> bool:
        """does the argument that starts here (up to its top-level `,` or `)`) call one of `names`"""
        depth = 0
        for t in self.t[self.i:]:
            if t == ("ch", "("): depth += 1
            elif t == ("ch", ")"):
                if depth == 0: return False
                depth -= 1
            elif t == ("ch", ",") and depth == 0: return False
            elif t[0] == "id" and t[1] in names: return True
        return False

    def statement(self) -> api.Property:
        ident = self.expect("id")[1]; self.expect("ch", "=")
        proc = self.expect("id")[1]; self.expect("ch", "(")
        self.arg_meta = []   # how each argument of distance / angle / dihedral was written (index()): decides its meaning inside `in` contexts
        if proc == "rdf":
            wr = None
            if self._has_within_before_comma():   # dynamic reference set: within([min:]max, selection), optionally `and` a static selection
                wlo, wr, wsel, wand = self.dyn_selection()
            grp = self.groups() if wr is None else None
            ref = None if (grp is not None or wr is not None) else self.selection()
            self.expect("ch", ",")
            trg = self.sel_or_within() if self._has_within_before_comma() else self.groups_or_selection()   # an array of selections as target: one centre of mass each
            self.expect("ch", ",")
            a = self.number(); lo, hi = 0.0, a
            if self.peek() == ("ch", ":"):
                self.next(); lo, hi = a, self.number()
            if wr is not None and isinstance(trg, list): raise ScriptError("a dynamic reference set with an array of selections as target is not lowered")
            if wr is not None: p = api.rdf(ident, api.Within(wr, wsel, wlo, wand), trg, hi, lo) if isinstance(trg, api.Within) else api.rdf_within(ident, wr, wsel, trg, hi, lo, wlo, wand)
            else: p = api.rdf_com(ident, grp, trg, hi, lo) if grp is not None else api.rdf(ident, ref, trg, hi, lo)
        elif proc == "sdf":
            st = self.structures(); self.expect("ch", ","); trg = self.sel_or_within(); self.expect("ch", ","); c = self.number()
            p = api.sdf(ident, st, trg, c)
        elif proc in ("density_x", "density_y", "density_z"):
            p = api.density(ident, "xyz".index(proc[-1]), self.sel_or_within())
        elif proc == "distance_pair":   # an array of selections (residue(a:b) over several residues) is one centre of mass per selection
            a = self.groups_or_selection(); self.expect("ch", ","); b = self.groups_or_selection()
            p = api.distance_pair(ident, a, b)
        elif proc in ("distance_min", "distance_max"):
            a = self.sel_or_within() if self._has_within_before_comma() else self.groups_or_selection(); self.expect("ch", ",")   # an array of selections: one centre of mass per selection
            b = self.sel_or_within() if self._has_within_before_comma() else self.groups_or_selection()
            p = {"distance_min": api.distance_min, "distance_max": api.distance_max}[proc](ident, a, b)
        elif proc == "contact_count":   # contact_count(A[], B, cutoff): the only registered signature (md_script_functions.inl:705); _contact_count's path length stays at its default 4 (:2762)
            a = self.groups_or_selection(); self.expect("ch", ","); b = self.selection(); self.expect("ch", ","); c = self.number()
            if self.peek() == ("ch", ","): raise ScriptError("Could not find matching procedure 'contact_count' which takes four arguments")
            p = api.contact_count(ident, a if isinstance(a, list) else [a], b, c, self.sys, 4)
        elif proc == "count":   # count(within(radius, selection)): the one dynamic selection the device path evaluates
            if not self._has_within_before_comma(): raise ScriptError("count() is lowered for within(radius, selection) expressions only")
            rlo, r, sel, cand = self.dyn_selection()
            p = api.count_within(ident, r, sel, rlo, cand)
        elif proc in ("coord_x", "coord_y", "coord_z"):
            a = self.index()   # an array of selections: one value per selection (its centre of mass, coordinate_extract :1503)
            p = api.coord(ident, "xyz".index(proc[-1]), a if isinstance(a, list) else ([a] if np.ndim(a) == 0 else a))
        elif proc == "com":
            p = api.com(ident, self.index())
        elif proc == "plane":
            p = api.plane(ident, self.groups_or_selection())   # an array of selections: the plane through their centres of mass
        elif proc == "rmsd":
            p = api.rmsd(ident, self.selection())   # an array of selections is flattened into their union (_internal_flatten_bf :4305)
        elif proc == "distance":
            a = self.index(True); self.expect("ch", ","); b = self.index(True); p = api.distance(ident, a, b)
        elif proc == "angle":
            a = self.index(); self.expect("ch", ","); b = self.index(); self.expect("ch", ","); c = self.index(); p = api.angle(ident, a, b, c)
        elif proc == "dihedral":
            v = [self.index()]
            for _ in range(3): self.expect("ch", ","); v.append(self.index())
            p = api.dihedral(ident, *v)
        else:
            raise ScriptError(f"procedure '{proc}' is outside the GPU hot-path scope")
        self.expect("ch", ")")
        if self.peek() == ("id", "in"):   # `expr in contexts` (evaluate_context md_script.c:3418): one value per context
            self.next(); begs, ends = self.contexts()
            if proc not in ("distance", "angle", "dihedral") or any(m[0] == "other" for m in self.arg_meta) or len(self.arg_meta) != len(p.idx):
                raise ScriptError("`in` is lowered for distance / angle / dihedral with integer or selection arguments")
            args = []
            for m in self.arg_meta:
                if m[0] == "int":   # remap_index_to_context rejects indices outside the context (md_script_functions.inl:1023-1040)
                    if np.any(begs + m[1] >= ends): raise ScriptError(f"supplied index ({m[1] + 1}) is not within the range of a context")
                    args.append(m[1])
                elif m[0] == "atomrange":   # atom(a:b) inside a context is relative to the context's first atom
                    if np.any(begs + m[2] > ends): raise ScriptError(f"supplied range ({m[1] + 1}:{m[2]}) is not within range of its context")
                    args.append([np.arange(b + m[1], b + m[2], dtype=np.int32) for b in begs])
                else:               # a selection: in context c the centre of mass of (selection AND context) (coordinate_extract_com with ctx->mol_ctx, :1812-1823)
                    sel = np.asarray(m[1], np.int64)
                    args.append([sel[(sel >= b) & (sel < e)].astype(np.int32) for b, e in zip(begs, ends)])
            p = api.in_contexts(ident, p.op, args, begs)
        self.expect("ch", ";")
        return p

    def contexts(self):
        """right-hand side of `in`: residue(a:b) | residue(:) | resname('X') -> (first atom, one past the last atom) of each context (md_bitfield beg_bit / end_bit)"""
        off = np.asarray(self.sys.res_atom_offset)
        f = self.expect("id")[1]; self.expect("ch", "(")
        if f == "residue":
            lo, hi = self._range(len(off) - 1); self.expect("ch", ")")
            return off[lo:hi].astype(np.int64), off[lo + 1:hi + 1].astype(np.int64)
        if f == "resname":
            if self.sys.resname is None: raise ScriptError("system has no residue data")
            pats = [self.expect("str")[1]]
            while self.peek() == ("ch", ","): self.next(); pats.append(self.expect("str")[1])
            self.expect("ch", ")")
            rn = np.asarray(self.sys.resname); hit = np.zeros(len(rn), bool)
            for pt in pats: hit |= np.array([fnmatch.fnmatchcase(r, pt) for r in rn])
            return off[:-1][hit].astype(np.int64), off[1:][hit].astype(np.int64)
        raise ScriptError(f"unsupported context expression '{f}'")


def compile_script(src: str, system: api.System) -> List[api.Property]:
    """`md_script_ir_compile_from_source` stand-in for the supported statement subset."""
    ps = _Parser(_tokens(src), system); out = []
    while ps.peek()[0] != "eof":
        out.append(ps.statement())
    if not out:
        raise ScriptError("No properties present in ir")
    return out
