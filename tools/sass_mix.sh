#!/bin/bash
# usage: tools/sass_mix.sh <kernel-name-substring>   -- opcode histogram of one kernel's SASS
cuobjdump -sass -fun "$1" lie_vae_b200/liblievae_sm100a.so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?([A-Z0-9_.]+).*/\2/' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -${2:-25}
