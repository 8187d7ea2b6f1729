"""On-disk SNVprofile layout written natively (the step after the hot path).

The reference's `inStrain.SNVprofile.SNVprofile` (SNVprofile.py:24-113,555-640,789-862) is a directory
    <ISP_loc>/{output,raw_data,log,figures}/   +   raw_data/attributes.tsv  (name, value, type, description)
whose attributes are files next to the table, one storage type per attribute: value (inline), dictionary (.json),
list (.txt), numpy (.npz), pandas (.csv.gz), pickle (.pickle) and `special` (covT / clonT -> .hd5, one gzip dataset
per "<scaffold>::<mm>").  The reference class needs h5py at import time; this module writes and reads the same layout
with instrain_b200.hd5, so `profile_bam` leaves an IS directory behind that `inStrain.SNVprofile.SNVprofile(ISP_loc)`
(compare, GeneProfile, polymorpher, plotting) opens unchanged.  Same method names and argument meaning: store / get.
"""
import json
import logging
import os
import pickle
import warnings

import numpy as np
import pandas as pd

from . import hd5

MIRRORED_VERSION = "1.9.1"          # inStrain/_version.py of the reference this layout follows
FIRST_LEVELS = ["output", "raw_data", "log", "figures"]
_README = ("The data in this folder can be easily accessed using the inStrain python API.\nFor information on how this is "
           "done, see the inStrain documentaion at https://instrain.readthedocs.io/en/latest/\n")
_EXT = {"dictionary": ".json", "list": ".txt", "numpy": ".npz", "pandas": ".csv.gz", "pickle": ".pickle", "special": ".hd5"}
_HD5_NAMES = ("covT", "clonT", "clonTR", "snpsCounted")


def _json_default(o):
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, np.floating):
        return float(o)
    if isinstance(o, (set, np.ndarray)):
        return list(o)
    raise TypeError(type(o))


class SNVprofileStore:
    def __init__(self, location):
        self.location = os.path.abspath(location)
        for lvl in FIRST_LEVELS:
            os.makedirs(os.path.join(self.location, lvl), exist_ok=True)
        if not os.path.exists(self._attributes_loc()):
            self._write_attributes(pd.DataFrame({"value": [], "type": [], "description": []}))
            self.store("location", self.location, "value", "Location of SNVprofile object")
            self.store("version", MIRRORED_VERSION, "value", "Version of inStrain")
            with open(self._fileloc("_README.txt"), "w") as o:
                o.write(_README)
        elif self.get("location") != self.location:
            self.store("location", self.location, "value", "Location of SNVprofile object")

    # ---- the reference's public pair -------------------------------------------------------------------------------
    def store(self, name, value, type, description):                      # noqa: A002 - the reference's argument name
        if type == "value":
            stored = value
        else:
            if type not in _EXT:
                logging.error("I dont know how to save a {0} type, so Im just going to pickle it".format(type))
                type = "pickle"
            if type == "special" and name not in _HD5_NAMES:
                logging.error("I dont know how to store {0}! Ill just pickle it".format(name))
                stored = self._fileloc(name) + ".pickle"
                self._save("pickle", value, stored)
            else:
                stored = self._fileloc(name) + _EXT[type]
                self._save(type, value, stored)
        Adb = self._read_attributes()
        if name in Adb.index:
            for thing, new in (("type", type), ("description", description)):
                if Adb.loc[name, thing] != new:
                    logging.error("WILL NOT OVERWRITE {0}; {1} arent the same ({2} vs {3}))".format(
                        name, thing, Adb.loc[name, thing], new))
                    return
            Adb.at[name, "value"] = stored
        else:
            Adb = pd.concat([Adb, pd.DataFrame({"value": stored, "type": type, "description": description}, index=[name])])
        self._write_attributes(Adb)

    def store_many(self, items, threads=4):
        """`store` for several (name, value, type, description) at once: the files are written concurrently (deflate of
        the csv.gz tables and of the .hd5 chunks releases the GIL), the attribute table is updated afterwards, in order.
        Same files, same attributes.tsv as the one-by-one calls."""
        from concurrent.futures import ThreadPoolExecutor
        plan = []
        for name, value, typ, description in items:
            if typ == "value":
                plan.append((name, value, typ, description, None, None))
                continue
            if typ not in _EXT:
                logging.error("I dont know how to save a {0} type, so Im just going to pickle it".format(typ))
                typ = "pickle"
            save_as = typ
            if typ == "special" and name not in _HD5_NAMES:
                logging.error("I dont know how to store {0}! Ill just pickle it".format(name))
                stored, save_as = self._fileloc(name) + ".pickle", "pickle"
            else:
                stored = self._fileloc(name) + _EXT[typ]
            plan.append((name, stored, typ, description, save_as, value))
        jobs = [e for e in plan if e[4] is not None]
        if threads > 1 and len(jobs) > 1:
            with ThreadPoolExecutor(max_workers=min(threads, len(jobs))) as ex:
                for f in [ex.submit(self._save, e[4], e[5], e[1]) for e in jobs]:
                    f.result()
        else:
            for e in jobs:
                self._save(e[4], e[5], e[1])
        Adb = self._read_attributes()
        for name, stored, typ, description, _, _ in plan:
            if name in Adb.index:
                clash = [t for t, new in (("type", typ), ("description", description)) if Adb.loc[name, t] != new]
                if clash:
                    logging.error("WILL NOT OVERWRITE {0}; {1} arent the same".format(name, clash[0]))
                    continue
                Adb.at[name, "value"] = stored
            else:
                Adb = pd.concat([Adb, pd.DataFrame({"value": stored, "type": typ, "description": description}, index=[name])])
        self._write_attributes(Adb)

    def get(self, name, **kwargs):
        Adb = self._read_attributes()
        if name not in Adb.index:
            return None
        typ = Adb.loc[name, "type"]
        if typ == "value":
            return Adb.loc[name, "value"]
        filename = os.path.join(self.location, "raw_data", os.path.basename(str(Adb.loc[name, "value"])))
        if typ == "dictionary":
            with open(filename) as fp:
                return json.load(fp)
        if typ == "list":
            with open(filename) as f:
                return [line.strip() for line in f]
        if typ == "numpy":
            return np.load(filename, allow_pickle=True)["arr_0"]
        if typ == "pandas":
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return pd.read_csv(filename, index_col=0)
        if typ == "special" and filename.endswith(".hd5"):
            return hd5.load_special(filename, scaffolds=kwargs.get("scaffolds", ()))
        if typ in ("pickle", "special"):
            with open(filename, "rb") as f:
                return pickle.load(f)
        logging.error("I dont know how to load a {0} type!".format(typ))
        return None

    def get_location(self, name):
        if name in FIRST_LEVELS:
            return os.path.join(self.location, name)
        raise KeyError(name)

    def __str__(self):
        return str(self._read_attributes())

    # ---- user-facing output tables (SNVprofile.generate, SNVprofile.py:192-442) ---------------------------------------------
    def get_output_base(self):
        return os.path.join(self.location, "output", os.path.basename(self.location) + "_")

    def _nonredundant(self, name, subset, drop_cryptic=False):
        """One row per `subset` key: the row of the highest mm (get_nonredundant_*_table, SNVprofile.py:483-522)."""
        db = self.get(name)
        if db is not None and drop_cryptic and "cryptic" in db:
            db = db[db["cryptic"] == False]                              # noqa: E712 - v1.6: cryptic SNVs are not reported
        if db is None or len(db) == 0:
            return pd.DataFrame()
        if "mm" not in db.columns:                                        # a table without levels is already nonredundant
            return db
        return db.sort_values("mm", kind="stable").drop_duplicates(subset=subset, keep="last").sort_index().drop(columns=["mm"])

    _OUTPUTS = {
        "SNVs": ("cumulative_snv_table", ["scaffold", "position"],
                 ["scaffold", "position", "position_coverage", "allele_count", "ref_base", "con_base", "var_base", "ref_freq",
                  "con_freq", "var_freq", "A", "C", "T", "G", "gene", "mutation", "mutation_type", "cryptic"]),
        "scaffold_info": ("cumulative_scaffold_table", ["scaffold"],
                          ["scaffold", "length", "coverage", "breadth", "nucl_diversity", "coverage_median", "coverage_std",
                           "coverage_SEM", "breadth_minCov", "breadth_expected", "nucl_diversity_median",
                           "nucl_diversity_rarefied", "nucl_diversity_rarefied_median", "breadth_rarefied", "conANI_reference",
                           "popANI_reference", "SNS_count", "SNV_count", "divergent_site_count"]),
        "linkage": ("raw_linkage_table", ["scaffold", "position_A", "position_B"],
                    ["scaffold", "position_A", "position_B", "distance", "r2", "d_prime", "r2_normalized", "d_prime_normalized",
                     "allele_A", "allele_a", "allele_B", "allele_b", "countab", "countAb", "countaB", "countAB", "total"]),
    }

    def generate(self, name, store=True, return_table=False, **kwargs):
        """The user-facing tables the hot path feeds -- SNVs, scaffold_info, linkage -- as the reference's
        SNVprofile.generate writes them into output/ (one row per site / scaffold / site pair at its highest mm, the
        reference's column order, .tsv, .tsv.gz beyond 1e6 rows).  gene_info / genome_info / mapping_info come from other
        modules of inStrain and are not produced here."""
        if name == "mapping_info":                                       # the read filter's report, with its settings as a header
            db = self.get("mapping_info")                                # line (write_mapping_info, filter_reads.py:699-720)
            if db is None:
                return None
            order = ["scaffold", "pass_pairing_filter", "filtered_pairs"]
            db = db[order + [c for c in db.columns if c not in order]]
            if store:
                values = {"min_read_ani": kwargs.get("min_read_ani", 0.97), "max_insert_relative": kwargs.get("max_insert_relative", 3),
                          "min_insert": kwargs.get("min_insert", 50), "min_mapq": kwargs.get("min_mapq", 2),
                          "pairing_filter": kwargs.get("pairing_filter", "paired_only")}
                ft = ".tsv.gz" if kwargs.get("force_compress", False) else ".tsv"
                with open(self.get_output_base() + name + ft, "w") as f:
                    f.write("# {0}\n".format(" ".join("{0}:{1}".format(k, v) for k, v in values.items())))
                    db.to_csv(f, index=False, sep="\t")
            return db if return_table else None
        if name == "gene_info":                                          # SNVprofile.py:247-257: nothing to write without a gene file
            if self.get("genes_table") is None:
                logging.info("Cannot generate genes_table, no genes were profiled")
            else:                                                         # GeneProfile (out of this path's scope) stored one
                logging.info("gene_info is written by inStrain's own SNVprofile.generate; open this directory with it")
            return None
        if name == "genome_info":                                        # produced by genomeUtilities, downstream of the path
            if self.get("genome_level_info") is None:
                logging.info("Cannot generate genome_info, no genome-level profile was stored")
            return None
        if name not in self._OUTPUTS:
            raise KeyError("generate(%r): only %s come out of the profile hot path" % (name, sorted(self._OUTPUTS) + ["mapping_info"]))
        source, subset, order = self._OUTPUTS[name]
        db = self._nonredundant(source, subset, drop_cryptic=(name == "SNVs"))
        if len(db) > 0:                                                   # reorder_columns (SNVprofile.py:1151-1165);
            cols = set(db.columns)                                        # columns outside the order keep the table's order
            db = db[[c for c in order if c in cols] + [c for c in db.columns if c not in order]]
        if store:
            ft = ".tsv.gz" if (len(db) > 1e6 or kwargs.get("force_compress", False)) else ".tsv"
            db.to_csv(self.get_output_base() + name + ft, index=False, sep="\t")
        return db if return_table else None

    # ---- storage back ends --------------------------------------------------------------------------------------------
    def _save(self, typ, value, loc):
        if typ == "dictionary":
            assert isinstance(value, dict)
            with open(loc, "w") as fp:
                json.dump(value, fp, default=_json_default)
        elif typ == "list":
            assert isinstance(value, list)
            with open(loc, "w") as f:
                for s in value:
                    f.write(str(s) + "\n")
        elif typ == "numpy":
            np.savez_compressed(loc, value)
        elif typ == "pandas":
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                # gzip level 4 instead of pandas' 9: 3 % larger files, half the time (the reference's reader is pd.read_csv)
                value.to_csv(loc, compression={"method": "gzip", "compresslevel": 4} if loc.endswith(".gz") else "infer")
        elif typ == "special":
            hd5.store_special(loc, value)
        else:
            with open(loc, "wb") as f:
                pickle.dump(value, f, pickle.HIGHEST_PROTOCOL)

    def _fileloc(self, name):
        return os.path.join(self.location, "raw_data/{0}".format(name))

    def _attributes_loc(self):
        return os.path.join(self.location, "raw_data/attributes.tsv")

    def _read_attributes(self):
        return pd.read_csv(self._attributes_loc(), sep="\t", index_col="name")

    def _write_attributes(self, Adb):
        Adb.to_csv(self._attributes_loc(), sep="\t", index_label="name")


class ProfileStore(SNVprofileStore):
    """What profile_bam returns when inStrain's own SNVprofile class is not importable: the on-disk object (store / get /
    generate / get_location, the interface ProfileController uses on `self.ISP`, controller.py:341-360) with the run's
    in-memory ProfileResult attached; attribute names the store does not have (scaffold_list, raw_snp_table, scaffolds,
    failures, timing ...) are answered by that result."""

    def __init__(self, location, result=None):
        super().__init__(location)
        self.result = result

    def __getattr__(self, name):                                        # only called when normal lookup fails
        res = self.__dict__.get("result")
        if res is not None and not name.startswith("__") and hasattr(res, name):
            return getattr(res, name)
        raise AttributeError(name)


def store_profile(ISP_loc, bam, res, mapping_info=None, **kwargs):
    """What gen_snv_profile stores for a profile run (profile_utilities.py:670-706), from a ProfileResult; plus the read
    filter's report when profile_bam ran the filter itself (ProfileController.load_paired_reads stores it,
    controller.py:301-304).  kwargs: the filter settings for the header of output/*_mapping_info.tsv."""
    S = ProfileStore(ISP_loc, res)
    if mapping_info is not None:
        S.store("mapping_info", mapping_info, "pandas", "Report on reads")
    items = [("object_type", "profile", "value", "Type of SNVprofile (profile or compare)"),
             ("bam_loc", bam, "value", "Location of .bam file"),
             ("scaffold_list", list(res.scaffold_list), "list", "1d list of scaffolds that were profiled"),
             ("raw_linkage_table", res.raw_linkage_table, "pandas", "Raw table of linkage information"),
             ("raw_snp_table", res.cumulative_snv_table, "pandas", "Contains raw SNP information on a mm level"),
             ("cumulative_scaffold_table", res.cumulative_scaffold_table, "pandas",
              "Cumulative coverage on mm level. Formerly scaffoldTable.csv"),
             ("cumulative_snv_table", res.cumulative_snv_table, "pandas", "Cumulative SNP on mm level. Formerly snpLocations.pickle"),
             ("scaffold_2_mm_2_read_2_snvs", {}, "pickle", "crazy nonsense needed for linkage"),
             ("covT", {s: p.covT for s, p in res.scaffolds.items()}, "special", "Scaffold -> mm -> position based coverage"),
             ("clonT", {s: p.clonT for s, p in res.scaffolds.items()}, "special", "Scaffold -> mm -> position based clonality")]
    if any(p.clonTR for p in res.scaffolds.values()):
        items.append(("clonTR", {s: p.clonTR for s, p in res.scaffolds.items()}, "special",
                      "Scaffold -> mm -> rarefied position based clonality"))
    if any(p.pileup_counts is not None for p in res.scaffolds.values()):   # --store_everything (profile_utilities.py:709-715)
        items.append(("counts_table", [res.scaffolds[s].pileup_counts for s in res.scaffold_list if s in res.scaffolds], "pickle",
                      "1d numpy array of 2D counts tables for each scaffold"))
    S.store_many(items, threads=int(kwargs.pop("store_threads", 0) or min(8, os.cpu_count() or 1)))
    for name in ("SNVs", "scaffold_info", "linkage"):                     # ProfileController.write_output (controller.py:352-360)
        S.generate(name)
    if mapping_info is not None:
        S.generate("mapping_info", **kwargs)
    return S
