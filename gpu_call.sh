for i in 19 20 21 22; do timeout 60 ./build/fa_bringup $i 2>&1 | grep -v "^device\|^bringup"; done
