"""
Generates tests/golden/example_geojson_rings.json from the reference's example scene
(/root/reference/examples/example.geojson, also symlinked as tests/example.geojson there).
Only the polygon rings (input DATA: OpenStreetMap coordinates) are kept — no reference source.
Run in the build container:  python tests/golden/make_geojson_fixture.py
"""
import json
import os

SRC = "/root/reference/examples/example.geojson"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_geojson_rings.json")

d = json.load(open(SRC))
rings = [f["geometry"]["coordinates"][0] for f in d["features"]
         if f.get("geometry") and f["geometry"]["type"] == "Polygon"]
json.dump({"source": "examples/example.geojson (reference v0.4.0)", "rings": rings}, open(DST, "w"), indent=0)
print(DST, [len(r) for r in rings])
