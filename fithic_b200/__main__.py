from .fithic import main

main()
