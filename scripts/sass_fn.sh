#!/bin/bash
# usage: sass_fn.sh lib.so mangled-name  -> instruction lines only
cuobjdump -sass -fun "$2" "$1" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
