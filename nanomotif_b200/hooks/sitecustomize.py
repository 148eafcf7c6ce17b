"""Start-up hook of worker processes (see nanomotif_b200/patch.py): with NMB200_PATCH=1 in the environment and this
directory on PYTHONPATH, every interpreter -- including the workers nanomotif spawns (find_motifs_bin.py:323) --
patches nanomotif's scoring operators with the B200 backend the moment nanomotif imports them."""
import os

if os.environ.get("NMB200_PATCH") == "1":
    try:
        from nanomotif_b200.patch import install_import_hook

        install_import_hook()
    except Exception as exc:  # never take an interpreter down from sitecustomize; the parent checks the patch
        import sys

        print(f"nanomotif_b200: worker hook not installed: {exc!r}", file=sys.stderr)
