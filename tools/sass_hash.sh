#!/bin/bash
# usage: tools/sass_hash.sh lib.so   -- compare before/after a refactor that must not change the generated code
# hash of the instruction streams only (no file paths, line tables or encodings)
cuobjdump -sass "$1" | grep -E "^\s+/\*[0-9a-f]{4,}\*/|Function :" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s*\/\*.*$//' | md5sum | cut -c1-12
