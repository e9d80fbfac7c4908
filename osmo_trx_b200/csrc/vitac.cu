// vitac.cu — placeholder, filled in below in this round
