#!/bin/bash
# Prove that a source change leaves the device code of every ahead-of-time shape class untouched (used at the end of
# round 1, without GPU access, for the flag-guarded first-layer hoist): build HEAD's sources and the working tree's,
# compare the instruction streams (cuobjdump -sass, addresses / encodings / symbol names stripped) object by object.
#   usage: tools/sass_identity.sh        (working tree must differ from HEAD only by the change under test)
set -e
norm() { cuobjdump -sass $1 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\*[0-9a-f]+\*/##g' | sed -E 's/\s+/ /g' | md5sum | cut -d' ' -f1; }
objs=$(python - <<'PY'
from neuromancer_b200 import build
import os
print(" ".join(sorted({os.path.splitext(os.path.basename(build.write_spec(p)))[0] for p in build.aot_plans()})))
PY
)
rm -rf /tmp/sass_new /tmp/sass_old; mkdir -p /tmp/sass_new /tmp/sass_old
python -m neuromancer_b200.build > /dev/null
for o in $objs nm_api; do norm neuromancer_b200/_C/obj/$o.o > /tmp/sass_new/$o; done
git stash -q; python -m neuromancer_b200.build > /dev/null
for o in $objs nm_api; do norm neuromancer_b200/_C/obj/$o.o > /tmp/sass_old/$o; done
git stash pop -q; python -m neuromancer_b200.build > /dev/null
n=0; bad=0
for o in $objs nm_api; do n=$((n+1)); if ! cmp -s /tmp/sass_old/$o /tmp/sass_new/$o; then echo "DIFFERENT device code: $o"; bad=$((bad+1)); fi; done
echo "compared $n objects, $bad differ"
