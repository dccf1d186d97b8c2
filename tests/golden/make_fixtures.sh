#!/bin/bash
# regenerates tests/golden/fixtures from the reference checkout (run in the build container, where /root/reference exists)
set -e
here=$(cd "$(dirname "$0")" && pwd)
ref=${1:-/root/reference/tests}
mkdir -p "$here/fixtures"
for f in bc_se.txt bc_pe_fa.txt bc_pe_fq.txt test.fq test.fq.gz test.fq.zst test.fa.gz reads_1.fa.gz reads_1.fa.bz2 reads_1.fa.xz reads_2.fa.gz; do
  cp "$ref/$f" "$here/fixtures/$f"
done
(cd "$here/fixtures" && sha256sum *)
