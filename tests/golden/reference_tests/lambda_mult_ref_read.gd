#=GENOME_DIFF	1.0
DEL	1	29	NC_001416-0	139	1
INS	2	30	NC_001416-1	4566	G
SNP	3	31	NC_001416-2	1261	G
INS	4	32	NC_001416-2	1435	C
SNP	5	33	NC_001416-2	2314	A
SNP	7	34	NC_001416-3	1915	C
SNP	8	35	NC_001416-3	5833	G
DEL	9	36	NC_001416-3	8717	1
SNP	10	37	NC_001416-4	6817	C
INS	11	38	NC_001416-4	8156	A
SNP	12	39	NC_001416-4	8184	T
SNP	13	40	NC_001416-4	8191	T
SNP	14	41	NC_001416-4	8203	A
SNP	15	42	NC_001416-4	8328	G
SNP	16	43	NC_001416-4	8342	T
SNP	17	44	NC_001416-4	8442	A
SNP	18	45	NC_001416-4	8514	A
SNP	19	46	NC_001416-4	8559	A
SNP	20	47	NC_001416-4	8597	T
SNP	21	48	NC_001416-4	8708	C
SNP	22	49	NC_001416-4	8728	T
SNP	23	50	NC_001416-4	8774	A
SNP	24	51	NC_001416-4	8868	C
SNP	25	52	NC_001416-4	9077	G
SNP	26	53	NC_001416-4	9172	C
SUB	27	54,55	NC_001416-4	9176	2	AC
SNP	28	56	NC_001416-4	9359	C
RA	29	.	NC_001416-0	139	0	G	.	allele_frequencies=.:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.276e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=.	major_cov=9/9	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=9/9	prediction=consensus	ref_cov=0/0	score=63.9	total_cov=9/9
RA	30	.	NC_001416-1	4566	1	.	G	allele_frequencies=G:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.510e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=G	major_cov=15/12	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=15/12	prediction=consensus	ref_cov=0/0	score=72.1	total_cov=15/12
RA	31	.	NC_001416-2	1261	0	A	G	allele_frequencies=G:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.450e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=G	major_cov=8/16	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=8/16	new_seq=G	prediction=consensus	ref_cov=0/0	ref_seq=A	score=62.7	total_cov=8/16
RA	32	.	NC_001416-2	1435	1	.	C	allele_frequencies=C:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.559e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=C	major_cov=9/21	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=9/21	new_seq=N	prediction=consensus	ref_cov=0/0	ref_seq=C	score=86.8	total_cov=9/21
RA	33	.	NC_001416-2	2314	0	G	A	allele_frequencies=A:9.470e-01,G:2.199e-02,.:3.104e-02	fisher_strand_p_value=3.22581e-01	frequency=9.470e-01	frequency_lower=8.511e-01	frequency_upper=9.942e-01	ks_quality_p_value=9.35484e-01	major_base=A	major_cov=9/21	major_frequency=9.470e-01	minor_base=.	minor_cov=1/0	new_cov=9/21	prediction=consensus	ref_cov=0/1	score=83.5	total_cov=10/22
RA	34	.	NC_001416-3	1915	0	T	C	allele_frequencies=C:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.647e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=C	major_cov=23/15	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=23/15	prediction=consensus	ref_cov=0/0	score=100.1	total_cov=24/15
RA	35	.	NC_001416-3	5833	0	A	G	allele_frequencies=G:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.372e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=G	major_cov=5/16	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=5/16	new_seq=G	prediction=consensus	ref_cov=0/0	ref_seq=A	score=51.0	total_cov=5/16
RA	36	.	NC_001416-3	8717	0	C	.	allele_frequencies=.:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.403e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=.	major_cov=15/7	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=15/7	prediction=consensus	ref_cov=0/0	score=78.7	total_cov=15/7
RA	37	.	NC_001416-4	6817	0	T	C	allele_frequencies=C:9.793e-01,T:2.068e-02	fisher_strand_p_value=1.00000e+00	frequency=9.793e-01	frequency_lower=9.249e-01	frequency_upper=9.984e-01	ks_quality_p_value=5.43478e-01	major_base=C	major_cov=23/22	major_frequency=9.793e-01	minor_base=T	minor_cov=0/1	new_cov=23/22	new_seq=C	prediction=consensus	ref_cov=0/1	ref_seq=T	score=123.2	total_cov=24/23
RA	38	.	NC_001416-4	8156	1	.	A	allele_frequencies=A:9.525e-01,.:4.746e-02	fisher_strand_p_value=1.00000e+00	frequency=9.525e-01	frequency_lower=8.371e-01	frequency_upper=9.950e-01	ks_quality_p_value=6.66667e-01	major_base=A	major_cov=9/11	major_frequency=9.525e-01	minor_base=.	minor_cov=0/1	new_cov=9/11	prediction=consensus	ref_cov=0/1	score=58.5	total_cov=9/12
RA	39	.	NC_001416-4	8184	0	C	T	allele_frequencies=T:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.598e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=T	major_cov=17/16	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=17/16	prediction=consensus	ref_cov=0/0	score=95.7	total_cov=17/16
RA	40	.	NC_001416-4	8191	0	C	T	allele_frequencies=T:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.492e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=T	major_cov=12/14	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=12/14	prediction=consensus	ref_cov=0/0	score=71.5	total_cov=12/14
RA	41	.	NC_001416-4	8203	0	G	A	allele_frequencies=A:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.468e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=A	major_cov=12/13	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=12/13	prediction=consensus	ref_cov=0/0	score=63.8	total_cov=12/13
RA	42	.	NC_001416-4	8328	0	A	G	allele_frequencies=G:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.510e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=G	major_cov=14/13	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=14/13	new_seq=G	prediction=consensus	ref_cov=0/0	ref_seq=A	score=73.8	total_cov=14/13
RA	43	.	NC_001416-4	8342	0	C	T	allele_frequencies=C:3.436e-02,T:9.591e-01	fisher_strand_p_value=1.00000e+00	frequency=9.591e-01	frequency_lower=8.444e-01	frequency_upper=9.991e-01	ks_quality_p_value=1.00000e+00	major_base=T	major_cov=15/10	major_frequency=9.591e-01	minor_base=C	minor_cov=1/0	new_cov=15/10	new_seq=T	prediction=consensus	ref_cov=1/0	ref_seq=C	score=69.5	total_cov=16/11
RA	44	.	NC_001416-4	8442	0	G	A	allele_frequencies=A:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.308e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=A	major_cov=6/13	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=6/13	new_seq=A	prediction=consensus	ref_cov=0/0	ref_seq=G	score=51.6	total_cov=6/13
RA	45	.	NC_001416-4	8514	0	G	A	allele_frequencies=A:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.641e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=A	major_cov=20/17	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=20/17	new_seq=A	prediction=consensus	ref_cov=0/0	ref_seq=G	score=109.1	total_cov=20/17
RA	46	.	NC_001416-4	8559	0	G	A	allele_frequencies=A:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.511e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=A	major_cov=10/17	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=10/17	new_seq=A	prediction=consensus	ref_cov=0/0	ref_seq=G	score=77.5	total_cov=10/17
RA	47	.	NC_001416-4	8597	0	C	T	allele_frequencies=T:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.450e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=T	major_cov=14/10	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=14/10	new_seq=T	prediction=consensus	ref_cov=0/0	ref_seq=C	score=66.7	total_cov=14/10
RA	48	.	NC_001416-4	8708	0	T	C	allele_frequencies=C:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.312e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=C	major_cov=12/7	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=12/7	new_seq=C	prediction=consensus	ref_cov=0/0	ref_seq=T	score=50.2	total_cov=12/7
RA	49	.	NC_001416-4	8728	0	C	T	allele_frequencies=T:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.403e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=T	major_cov=8/14	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=8/14	new_seq=T	prediction=consensus	ref_cov=0/0	ref_seq=C	score=62.2	total_cov=8/14
RA	50	.	NC_001416-4	8774	0	C	A	allele_frequencies=A:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.358e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=A	major_cov=10/17	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=10/17	new_seq=A	prediction=consensus	ref_cov=1/0	ref_seq=C	score=74.0	total_cov=11/17
RA	51	.	NC_001416-4	8868	0	T	C	allele_frequencies=C:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.630e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=C	major_cov=20/16	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=20/16	prediction=consensus	ref_cov=0/0	score=97.0	total_cov=20/16
RA	52	.	NC_001416-4	9077	0	A	G	allele_frequencies=G:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.491e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=G	major_cov=11/15	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=11/15	new_seq=G	prediction=consensus	ref_cov=0/0	ref_seq=A	score=68.4	total_cov=11/15
RA	53	.	NC_001416-4	9172	0	T	C	allele_frequencies=C:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.558e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=C	major_cov=15/15	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=15/15	prediction=consensus	ref_cov=0/0	score=80.3	total_cov=15/15
RA	54	.	NC_001416-4	9176	0	G	A	allele_frequencies=A:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.527e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=A	major_cov=14/14	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=14/14	prediction=consensus	ref_cov=0/0	score=79.8	total_cov=14/14
RA	55	.	NC_001416-4	9177	0	T	C	allele_frequencies=C:1.000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.000e+00	frequency_lower=9.528e-01	frequency_upper=1.000e+00	ks_quality_p_value=1.00000e+00	major_base=C	major_cov=14/14	major_frequency=1.000e+00	minor_base=N	minor_cov=0/0	new_cov=14/14	prediction=consensus	ref_cov=0/0	score=76.8	total_cov=14/14
RA	56	.	NC_001416-4	9359	0	T	C	allele_frequencies=C:8.334e-01,T:1.666e-01	fisher_strand_p_value=6.22079e-01	frequency=8.334e-01	frequency_lower=7.032e-01	frequency_upper=9.247e-01	ks_quality_p_value=8.28597e-01	major_base=C	major_cov=11/14	major_frequency=8.334e-01	minor_base=T	minor_cov=1/4	new_cov=11/14	prediction=consensus	ref_cov=1/4	score=61.0	total_cov=12/18
MC	57	.	NC_001416-0	1	2	0	0	gene_name=–/nu1	gene_position=intergenic (–/-189)	gene_product=–/DNA packaging protein	gene_strand=–/>	left_inside_cov=0	left_outside_cov=NA	locus_tag=–/lambdap01	right_inside_cov=0	right_outside_cov=47
MC	58	.	NC_001416-2	2338	8333	0	0	gene_name=[orf-314]–ea59	gene_product=[orf-314],orf-194,ea47,ea31,ea59	left_inside_cov=0	left_outside_cov=26	locus_tag=[lambdap28]–[lambdap82]	right_inside_cov=9	right_outside_cov=36
