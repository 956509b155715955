#!/bin/bash
cd "$(dirname "$0")/bin"
for f in psd_*; do ./$f 2097152 17; done
