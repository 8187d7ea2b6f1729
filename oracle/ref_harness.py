"""Drive the reference's OWN hot-path functions (oracle / test infrastructure only).

Works only where /root/reference exists (the build container).  The six third-party packages the
reference imports at module import time but that are absent here (pysam, h5py, Bio, seaborn,
matplotlib, lmfit) are replaced by permissive stub modules, so that

    inStrain.profile.profile_utilities.process_bam_sites   (profile_utilities.py:218-266)
    inStrain.profile.snv_utilities.update_snp_table         (snv_utilities.py:40-145)
    inStrain.profile.linkage.calc_mm_SNV_linkage_network    (linkage.py:14-44)
    inStrain.profile.linkage.calculate_ld                   (linkage.py:46-75)

run UNMODIFIED, fed with duck-typed pileup columns built from event arrays.  Used to
(1) validate oracle/pileup_emul.py against the reference's golden tables and
(2) generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py).
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from collections import defaultdict

import numpy as np

REFERENCE_ROOT = os.environ.get("INSTRAIN_REFERENCE", "/root/reference")


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return _Stub("x")

    def __setitem__(self, k, v):
        pass

    def __getitem__(self, k):
        return _Stub("x")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    roots = {"pysam", "h5py", "Bio", "seaborn", "matplotlib", "lmfit"}

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, m):
        pass


_loaded = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "inStrain"))


def load_reference():
    """Import the reference's hot-path modules (pu, su, lk)."""
    global _loaded
    if _loaded is None:
        if not available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        sys.dont_write_bytecode = True
        for root in _Finder.roots:
            try:
                __import__(root)
            except Exception:
                pass
        missing = {r for r in _Finder.roots if r not in sys.modules}
        if missing:
            f = _Finder()
            f.roots = missing
            sys.meta_path.insert(0, f)
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import inStrain.profile.profile_utilities as pu
        import inStrain.profile.snv_utilities as su
        import inStrain.profile.linkage as lk
        _loaded = (pu, su, lk)
    return _loaded


def null_model(fdr=1e-6):
    """The reference's own null-model dict (snv_utilities.py:14-38)."""
    _, su, _ = load_reference()
    return su.generate_snp_model(os.path.join(REFERENCE_ROOT, "inStrain", "helper_files", "NullModel.txt"), fdr=fdr)


class _Aln:
    __slots__ = ("query_name", "query_sequence")

    def __init__(self, n, s):
        self.query_name, self.query_sequence = n, s


class _PRead:
    __slots__ = ("alignment", "query_position", "is_del", "is_refskip")

    def __init__(self, aln):
        self.alignment, self.query_position, self.is_del, self.is_refskip = aln, 0, False, False


class _Col:
    __slots__ = ("pos", "pileups")

    def __init__(self, pos, pileups):
        self.pos, self.pileups = pos, pileups


def columns_from_events(ev, start, end, min_qual=30):
    """Duck-typed pysam pileup columns of split [start, end] from event arrays (file order kept)."""
    keep = (ev["qual"] >= min_qual) & (ev["ref_pos"] >= start) & (ev["ref_pos"] <= end)
    idx = np.nonzero(keep)[0]
    pos = ev["ref_pos"][idx]
    order = np.argsort(pos, kind="stable")
    idx, pos = idx[order], pos[order]
    names = ev["names"]
    chars = "ACTGN"
    cols = []
    i, n = 0, len(idx)
    while i < n:
        j = i
        p = int(pos[i])
        pile = []
        while j < n and pos[j] == p:
            e = idx[j]
            pile.append(_PRead(_Aln(names[ev["read_id"][e]], chars[ev["base"][e]])))
            j += 1
        cols.append(_Col(p, pile))
        i = j
    return cols


def run_split(ev, seq, start, end, r2m, model, scaffold="s", min_cov=5, min_freq=0.05, min_snp=20,
              rarefied_coverage=50):
    """Reference `profile_split` body (profile_utilities.py:158-193) on emulated columns.

    Returns dict(snp=list of row dicts, ld=list of row dicts, covT={mm: int array[L]},
                 clonT={mm: float32 array[L]}) with positions ABSOLUTE (start added back, :181,188).
    generate_snp_table (snv_utilities.py:274-290) is bypassed: it raises on pandas 3 (:283-284);
    the table is assembled from Stable + p2c directly, which is what that function does.
    """
    pu, su, lk = load_reference()
    sseq = seq[start:end + 1]
    m_len = len(sseq)
    covT, clonT, clonTR, p2c = {}, {}, {}, {}
    read_to_snvs = defaultdict(pu._dlist)
    snv2mm2counts = {}
    stable = defaultdict(list)
    cols = columns_from_events(ev, start, end)
    pu.process_bam_sites(scaffold, sseq, iter(cols), covT, clonT, clonTR, p2c, read_to_snvs, snv2mm2counts,
                         stable, None, m_len, model, r2m, start=start, min_cov=min_cov, min_freq=min_freq,
                         min_snp=min_snp, rarefied_coverage=rarefied_coverage)
    snp_rows = []
    for i in range(len(stable["position"])):
        pos = stable["position"][i]
        snp_rows.append(dict(
            scaffold=scaffold, position=int(pos) + start, ref_base=stable["ref_base"][i],
            A=int(stable["A"][i]), C=int(stable["C"][i]), T=int(stable["T"][i]), G=int(stable["G"][i]),
            con_base=stable["con_base"][i], var_base=stable["var_base"][i], mm=int(stable["mm"][i]),
            allele_count=int(stable["allele_count"][i]), **{"class": stable["class"][i]},
            cryptic=bool(p2c.get(pos, False))))
    G = lk.calc_mm_SNV_linkage_network(read_to_snvs, scaff=scaffold)
    ld = lk.calculate_ld(G, min_snp, snv2mm2counts=snv2mm2counts, scaffold=scaffold)
    ld_rows = []
    if len(ld) > 0:
        for row in ld.to_dict("records"):
            row["position_A"] = int(row["position_A"]) + start
            row["position_B"] = int(row["position_B"]) + start
            ld_rows.append(row)
    return dict(snp=snp_rows, ld=ld_rows, covT=covT, clonT=clonT)
