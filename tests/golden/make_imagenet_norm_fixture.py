"""Extracts the reference's OWN golden vectors for ImageNetNormalization and ResizingMinMax from
/root/reference/test_units/augmentations/test_image_augmentations.py (IMG :5-15, the three float32
targets :21-64, the four output shapes :66-80) into tests/golden/imagenet_norm_ref.json.  The file is
parsed with ``ast`` -- TensorFlow is not needed (and not installable here).  Run in the build
container only; the GPU box uses the committed JSON.

    python tests/golden/make_imagenet_norm_fixture.py
"""
import ast
import json
import os

SRC = "/root/reference/test_units/augmentations/test_image_augmentations.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def first_list(node):
    """The first list literal inside a call like tf.cast([[...]], tf.float32)."""
    for n in ast.walk(node):
        if isinstance(n, ast.List):
            return ast.literal_eval(n)
    raise ValueError("no list literal")


def main():
    tree = ast.parse(open(SRC).read())
    out = {"source": "test_units/augmentations/test_image_augmentations.py", "targets": {}, "shapes": {}}
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "IMG" and "image" not in out:
            out["image"] = first_list(node.value)  # 4 x 4, stacked into 3 equal channels at :14
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef)][0]
    for fn in cls.body:
        if not isinstance(fn, ast.FunctionDef):
            continue
        if fn.name.startswith("test_imagenet_normalization_"):
            mode = fn.name.rsplit("_", 1)[1]
            assign = [n for n in fn.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == "target"][0]
            out["targets"][mode] = first_list(assign.value)
        elif fn.name.startswith("test_resize_"):
            kwargs, shape = None, None
            for n in ast.walk(fn):
                if isinstance(n, ast.Call) and getattr(n.func, "attr", None) == "ResizingMinMax":
                    kwargs = {k.arg: ast.literal_eval(k.value) for k in n.keywords}
                if isinstance(n, ast.Call) and getattr(n.func, "attr", None) == "assertAllEqual":
                    shape = ast.literal_eval(n.args[1])
            out["shapes"][fn.name] = {"kwargs": kwargs, "input_shape": [1, 4, 3, 3], "output_shape": shape}
    assert set(out["targets"]) == {"caffe", "tf", "torch"} and len(out["shapes"]) == 4
    with open(os.path.join(HERE, "imagenet_norm_ref.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out)[:300], "...")


if __name__ == "__main__":
    main()
