* max x + 4y + 8z  s.t.  3x + y <= 8,  y + 2z <= 7   (fixed-width MPS, two entries per line)
NAME          simple
ROWS
 N  obj
 L  row1
 L  row2
COLUMNS
    X         obj       1               row1      3
    Y         obj       4               row1      1
    Y         row2      1
    Z         obj       8               row2      2
RHS
    rhs1      row1      8               row2      7
ENDATA
