#!/bin/bash
# usage: tools/sass_stats.sh <mangled kernel substring>  -- load/store mix, stack and instruction count of one kernel
SO=/root/repo/motion-planning-for-autonomous-driving-with-mpc_b200/csrc/libmpcb200.so
mkdir -p /tmp/cub && cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all $SO >/dev/null && nvdisasm -g -c *.cubin > dis.txt 2>/dev/null
awk -v k="$1" '$0 ~ "^.text." && index($0,k){f=1} f' dis.txt | awk '/^\/\/-------/{if(NR>1)exit} {print}' > k.txt
echo "generic LD/ST: $(grep -c 'LD\.E\|ST\.E' k.txt)  LDS/STS: $(grep -c 'LDS\|STS' k.txt)  LDL/STL: $(grep -c 'LDL\|STL' k.txt)  SHFL: $(grep -c SHFL k.txt) MUFU: $(grep -c MUFU k.txt) total: $(grep -c '^\s*/\*[0-9a-f]*\*/' k.txt)"
