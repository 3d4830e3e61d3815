"""Mint tests/golden/ref_table_headers.json: YAML table headers in astropy's serialisation
format (the content of ``samples.__table_column_meta__`` in the reference's HDF5 prior
cache, thejoker/samples.py:535-563, utils.py:75-88).

* ``ecsv:<file>``: the header lines of the reference's own docs/examples/*.ecsv, written
  by astropy's table serialiser from QTables with Quantity columns and a Time in the meta
  -- the same ``get_yaml_from_table`` output that write_table_hdf5(serialize_meta=True)
  stores line by line in ``__table_column_meta__`` (the '# ' ECSV prefix removed).
* ``prior_samples``: a header for a prior-samples table [P, e, omega, M0, s, ln_prior]
  CONSTRUCTED here after the layout of those real headers (astropy, h5py and PyTables are
  not in the build image, so the reference cannot write one): units both in the datatype
  entries and in __serialized_columns__, t_ref / poly_trend / n_offsets in the !!omap meta.
Run in the build container (needs /root/reference)."""
import json
import os

REF = "/root/reference/docs/examples"
out = {}
for name in ("data.ecsv", "data-survey1.ecsv", "data-triple.ecsv"):
    lines = []
    for ln in open(os.path.join(REF, name)):
        if not ln.startswith("#"):
            break
        ln = ln[2:].rstrip("\n")
        if ln.startswith("%ECSV") or ln == "---":
            continue
        lines.append(ln)
    out["ecsv:" + name] = lines


def ser(name, unit):
    return [f"    {name}:", "      __class__: astropy.units.quantity.Quantity",
            f"      unit: !astropy.units.Unit {{unit: {unit}}}",
            f"      value: !astropy.table.SerializedColumn {{name: {name}}}"]


cols = [("P", "d"), ("e", None), ("omega", "rad"), ("M0", "rad"), ("s", "m / s"), ("ln_prior", None)]
hdr = ["datatype:"]
for n, un in cols:
    hdr.append(f"- {{name: {n}, unit: {un}, datatype: float64}}" if un else
               f"- {{name: {n}, datatype: float64}}")
hdr += ["meta: !!omap", "- {t_ref: null}", "- {poly_trend: 1}", "- {n_offsets: 0}",
        "- __serialized_columns__:"]
for n, un in cols:
    if un:
        hdr += ser(n, un)
hdr.append("schema: astropy-2.0")
out["prior_samples"] = hdr
# the same without units in the datatype entries (only the serialised-column section has them)
out["prior_samples_units_in_meta_only"] = [
    ln if not ln.startswith("- {name:") else
    "- {name: " + ln.split("name: ")[1].split(",")[0].rstrip("}") + ", datatype: float64}" for ln in hdr]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                 "ref_table_headers.json"), "w"), indent=1)
print({k: len(v) for k, v in out.items()})
